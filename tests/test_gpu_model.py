"""The executable specification of the kernels' index arithmetic (tests/gpu_model.py) must agree
byte for byte with the oracle -- on the reference's goldens and on adversarial random FASTA, at
tile sizes small enough that every tile-boundary case occurs."""
import pytest

import oracle
import gpu_model
from fasta_gen import mid_fasta, nasty_fasta, random_options
from optparse_ref import parse_option_string

MASKS = {"1": 1, "2": 2, "3": 4, "F": 7, "-1": 8, "-2": 16, "-3": 32, "R": 56, "6": 63}


def model(data, frame="1", table=0, clean=False, alternative=False, trim=False, tile=48, residues="windows"):
    aa64 = oracle.table_code(table)
    return gpu_model.translate(data, aa64, MASKS[frame], clean, alternative, trim, tile, residues)


@pytest.mark.parametrize("idx", range(29))
def test_model_on_reference_goldens(goldens, idx):
    case = goldens["cases"][idx]
    kw = parse_option_string(case["options"])
    for tile in (12, 48, 12288):
        assert model(goldens["input"].encode(), tile=tile, **kw).decode() == case["expected"]


@pytest.mark.parametrize("seed", range(150))
def test_model_vs_oracle_random(seed):
    data = nasty_fasta(seed, max_records=6, max_len=200)
    kw = random_options(seed)
    want = oracle.translate(data, **kw)
    for tile in (3, 12, 48, 300):
        assert model(data, tile=tile, **kw) == want, (seed, tile, kw)


def test_model_quirks():
    for data in (b"", b"\n", b"\r", b"\r\n", b">", b">\n", b">\r", b"A", b"\n\nAC", b">a\n>b\n>c",
                 b"AC\r", b">a b\r\nAC\r\r\nG", b"x>y\n>z"):
        for frame in ("6", "F", "R", "2"):
            for trim in (False, True):
                assert model(data, frame=frame, trim=trim, tile=3) == oracle.translate(data, frame=frame, trim=trim), (data, frame, trim)


# ---- the k_lines formulation: blocks of 180 nucleotides owned by the tile that holds their anchor ----

@pytest.mark.parametrize("idx", range(29))
def test_lines_model_on_reference_goldens(goldens, idx):
    case = goldens["cases"][idx]
    kw = parse_option_string(case["options"])
    for tile in (12, 48, 12288):
        assert model(goldens["input"].encode(), tile=tile, residues="lines", **kw).decode() == case["expected"]


@pytest.mark.parametrize("seed", range(60))
def test_lines_model_vs_oracle_random(seed):
    data = nasty_fasta(seed, max_records=6, max_len=200)
    kw = random_options(seed)
    want = oracle.translate(data, **kw)
    for tile in (3, 48, 300):
        assert model(data, tile=tile, residues="lines", **kw) == want, (seed, tile, kw)


@pytest.mark.parametrize("seed", range(24))
def test_lines_model_block_and_tile_edges(seed):
    """Records a few blocks long, tiles of a fraction of a block up to several blocks: every
    position of a block anchor relative to a tile edge and to the record's ends occurs."""
    import random
    rng = random.Random(1000 + seed)
    parts = []
    for r in range(rng.randint(1, 4)):
        n = rng.choice((0, 1, 2, 3, 59, 60, 61, 178, 179, 180, 181, 182, 187, 188, 189, 190, 191, 359, 360, 361,
                        rng.randint(0, 900)))
        seq = bytes(rng.choice(b"ACGTN") for _ in range(n))
        w = rng.choice((60, 7, 10 ** 6))
        parts.append(b">r%d x\n" % r + b"\n".join(seq[i:i + w] for i in range(0, n, w)) + b"\n")
    data = b"".join(parts)
    for frame in ("6", "F", "R", rng.choice(("1", "2", "3", "-1", "-2", "-3"))):
        for alt in (False, True):
            trim = rng.random() < 0.5
            want = oracle.translate(data, frame=frame, alternative=alt, trim=trim)
            for tile in (3, 45, 180, 183, 600):
                got = model(data, frame=frame, alternative=alt, trim=trim, tile=tile, residues="lines")
                assert got == want, (seed, frame, alt, trim, tile)


def test_lines_model_quirks():
    for data in (b"", b"\n", b"\r", b">", b">\n", b"A", b"\n\nAC", b">a\n>b\n>c", b"AC\r", b">a b\r\nAC\r\r\nG", b"x>y\n>z"):
        for frame in ("6", "F", "R", "2"):
            for trim in (False, True):
                got = model(data, frame=frame, trim=trim, tile=3, residues="lines")
                assert got == oracle.translate(data, frame=frame, trim=trim), (data, frame, trim)


@pytest.mark.parametrize("seed", range(8))
def test_lines_model_mid_size(seed):
    """Several kilobytes with long headers, many short records and a few long ones, tiles of 8 KB like the kernel's."""
    data = mid_fasta(7000 + seed, target=12_000)
    kw = random_options(7000 + seed)
    want = oracle.translate(data, **kw)
    for tile in (8256, 600):
        assert model(data, tile=tile, residues="lines", **kw) == want, (seed, tile, kw)
