"""Regenerate tests/golden/emboss_goldens.json from the reference checkout.

The reference pins its translation path with one table-driven test
(transeq/transeq_test.go:14-62): the FASTA fixture transeq/testdata/test.fna is translated
with each option string in transeq/testdata/data.json and compared byte-for-byte with the
EMBOSS-generated `expected` text (transeq/testdata/generateTestData.sh:8-52).  /root/reference
does not exist on the GPU box, so both are folded into one committed JSON fixture:

    {"source": ..., "input": "<test.fna text>", "cases": [{"options": ..., "expected": ...}, ...]}

Run:  python tests/golden/make_goldens.py [/root/reference]
"""
import json
import os
import sys


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    td = os.path.join(ref, "transeq", "testdata")
    with open(os.path.join(td, "test.fna"), "rb") as f:
        fasta = f.read().decode("ascii")
    with open(os.path.join(td, "data.json"), "r") as f:
        cases = json.load(f)
    doc = {
        "source": "feliixx/gotranseq transeq/testdata/{test.fna,data.json} (EMBOSS transeq output)",
        "input": fasta,
        "cases": [{"options": c["options"], "expected": c["expected"]} for c in cases],
    }
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emboss_goldens.json")
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    print(f"wrote {out}: {len(doc['cases'])} cases, input {len(fasta)} bytes")


if __name__ == "__main__":
    main()
