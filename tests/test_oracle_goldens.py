"""Pins the CPU oracle (oracle/transeq_oracle.c) on the reference's own golden vectors:
transeq/testdata/test.fna x transeq/testdata/data.json (29 EMBOSS-generated cases, asserted by the
reference in transeq/transeq_test.go:55 with NumWorker = 1)."""
import pytest

import oracle
from optparse_ref import parse_option_string


def test_golden_count(goldens):
    assert len(goldens["cases"]) == 29
    assert len(goldens["input"]) == 1985


@pytest.mark.parametrize("idx", range(29))
def test_oracle_matches_reference_goldens(goldens, idx):
    case = goldens["cases"][idx]
    kw = parse_option_string(case["options"])
    out = oracle.translate(goldens["input"].encode("ascii"), num_worker=1, **kw)
    assert out.decode("ascii") == case["expected"], case["options"]


def test_oracle_errors_match_reference_texts():
    # transeq.go:107 and ncbicode/code.go:378
    with pytest.raises(oracle.OracleError, match=r"wrong value for -f \| --frame parameter: 7"):
        oracle.translate(b">a\nACG\n", frame="7")
    with pytest.raises(oracle.OracleError, match=r"invalid table code: 1$"):
        oracle.translate(b">a\nACG\n", frame="1", table=1)


def test_oracle_multiworker_is_a_permutation_of_single_worker(goldens):
    # transeq.go:135-159: with several workers the record order is scheduling dependent
    data = goldens["input"].encode("ascii") * 50
    one = oracle.translate(data, frame="6", num_worker=1)
    many = oracle.translate(data, frame="6", num_worker=4)
    assert len(one) == len(many)
    assert sorted(one.split(b">")) == sorted(many.split(b">"))


def test_oracle_quirks():
    t = lambda d, **k: oracle.translate(d, **k)
    # empty input still emits one record with an empty header (transeq.go:213)
    assert t(b"", frame="F") == b"_1\n_2\n_3\n"
    # bytes before the first '>' form a record with an empty header (transeq.go:198-210)
    assert t(b"ACGT\n>x\nACG", frame="1") == b"_1\nTX\n>x_1\nT\n"
    # _N goes before the first space only (writer.go:121)
    assert t(b">a b c\nATG", frame="1") == b">a_1 b c\nM\n"
    # one trailing CR per line is dropped (bufio.ScanLines); header-only record
    assert t(b">a\r\nATG\r\n>b\r\n", frame="1") == b">a_1\nM\n>b_1\n"
    # Go's % truncates: n=1 with start 2 emits nothing; n - start == 1 gives 'X'
    assert t(b">a\nA", frame="F") == b">a_1\nX\n>a_2\n>a_3\n"
    # two-letter codon at the tail (transeq.go:65-83): GTN -> V, ATN -> X
    assert t(b">a\nGT\n>b\nAT\n", frame="1") == b">a_1\nV\n>b_1\nX\n"
