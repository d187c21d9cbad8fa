"""Regenerate tests/golden/ncbi_tables.json from the reference checkout: the 21 EFFECTIVE codon
maps LoadTableCode returns (ncbicode/code.go:367-388) -- the standard map (code.go:16-81) overlaid
with the table's diff map (code.go:83-313, looked up through `diffs`, code.go:315-337, and the id
constants, code.go:341-363), U read as T in the diff keys (code.go:383).  Parsed from the Go source
text; /root/reference does not exist on the GPU box, so the result is a committed fixture:

    {"<id>": "<64 amino acids in T,C,A,G codon order, index 16*b0 + 4*b1 + b2>", ...}

Run:  python tests/golden/make_ncbi_tables.py [/root/reference]
"""
import json
import os
import re
import sys


def parse(src):
    maps = {}
    for m in re.finditer(r"(\w+)\s*=\s*map\[string\]byte\{(.*?)\}", src, flags=re.S):
        maps[m.group(1)] = dict(re.findall(r'"([ACGTU]{3})":\s*\'(.)\'', m.group(2)))
    ids = {name: int(v) for name, v in re.findall(r"^\s*(\w+)\s*=\s*(\d+)\s*$", src[src.index("const ("):], flags=re.M)}
    body = src[src.index("diffs = map[int]map[string]byte{"):]
    body = body[:body.index("\n\t}")]
    diff_of = dict(re.findall(r"^\s*(\w+):\s*(\w+),", body, flags=re.M))
    std = maps["standard"]
    assert len(std) == 64
    order = "TCAG"
    tables = {}
    for name, code in ids.items():
        eff = dict(std)
        if code != 0:
            for codon, aa in maps[diff_of[name]].items():
                eff[codon.replace("U", "T")] = aa
        assert len(eff) == 64, name
        tables[str(code)] = "".join(eff[a + b + c] for a in order for b in order for c in order)
    return tables


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    with open(os.path.join(ref, "ncbicode", "code.go")) as f:
        tables = parse(f.read())
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncbi_tables.json")
    with open(out, "w") as f:
        json.dump(dict(sorted(tables.items(), key=lambda kv: int(kv[0]))), f, indent=1)
        f.write("\n")
    print(f"wrote {out}: {len(tables)} tables")


if __name__ == "__main__":
    main()
