#!/usr/bin/env python3
"""Extract the reference's field known-answer vectors into tests/golden/field_kat.json.

Run HERE (build container) where /root/reference exists; the GPU box only sees the JSON.
Sources (relative to /root/reference):
  crates/field/src/binary_field.rs:925-1028   test_bin{2,4,8,16,64}b_mul assert_eq! vectors
  crates/field/src/binary_field.rs:740-747    MULTIPLICATIVE_GENERATOR of B1..B128
  crates/field/src/aes_field.rs:46-50, 113-141 AES tower generators + tower<->AES byte maps
  crates/field/src/polyval.rs:262, 496, 516-788, 1113-1127  POLYVAL ONE / generator / basis-change
                                               tables / mul + square KATs
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "field_kat.json")


def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


def main():
    bf = read("crates/field/src/binary_field.rs")
    kats = {}
    for bits, ty in [(2, "BF2"), (4, "BF4"), (8, "BF8"), (16, "BF16"), (64, "BF64")]:
        pat = re.compile(
            r"assert_eq!\(\s*%s::(?:from|new)\((0x[0-9a-fA-F]+)\)\s*\*\s*%s::(?:from|new)\((0x[0-9a-fA-F]+)\),\s*%s::(?:from|new)\((0x[0-9a-fA-F]+)\)\s*\)"
            % (ty, ty, ty), re.S)
        kats[str(bits)] = [[int(a, 16), int(b, 16), int(c, 16)] for a, b, c in pat.findall(bf)]
        assert len(kats[str(bits)]) >= 10, (bits, len(kats[str(bits)]))
    gens = {}
    for m in re.finditer(r"binary_field!\(pub BinaryField(\d+)b\(\w+\), (?:U\d+::new\()?(0x[0-9a-fA-F]+)", bf):
        gens[m.group(1)] = int(m.group(2), 16)
    assert set(gens) == {"1", "2", "4", "8", "16", "32", "64", "128"}, gens

    aes = read("crates/field/src/aes_field.rs")
    aes_gens = {m.group(1): int(m.group(2), 16)
                for m in re.finditer(r"binary_field!\(pub AESTowerField(\d+)b\(\w+\), (0x[0-9a-fA-F]+)\)", aes)}

    def table(src, name, ty):
        i = src.index("pub const " + name)
        j = src.index("]);", i)
        return [int(x, 16) for x in re.findall(r"%s\((0x[0-9a-fA-F]+)\)" % ty, src[i:j])]

    aes_to_bin = table(aes, "AES_TO_BINARY_LINEAR_TRANSFORMATION", "BinaryField8b")
    bin_to_aes = table(aes, "BINARY_TO_AES_LINEAR_TRANSFORMATION", "AESTowerField8b")
    assert len(aes_to_bin) == 8 and len(bin_to_aes) == 8

    pv = read("crates/field/src/polyval.rs")
    b2p = table(pv, "BINARY_TO_POLYVAL_TRANSFORMATION", "BinaryField128bPolyval")
    p2b = table(pv, "POLYVAL_TO_BINARY_TRANSFORMATION", "BinaryField128b")
    assert len(b2p) == 128 and len(p2b) == 128
    m = re.search(r"fn test_mul\(\).*?new\((0x[0-9a-f]+)\)\s*\*\s*BinaryField128bPolyval::new\((0x[0-9a-f]+)\),\s*"
                  r"BinaryField128bPolyval::new\((0x[0-9a-f]+)\)", pv, re.S)
    mul_kat = [int(x, 16) for x in m.groups()]
    m = re.search(r"fn test_sqr\(\).*?new\((0x[0-9a-f]+)\)\),\s*BinaryField128bPolyval::new\((0x[0-9a-f]+)\)", pv, re.S)
    sqr_kat = [int(x, 16) for x in m.groups()]
    one = int(re.search(r"const ONE: Self = Self\((0x[0-9a-f]+)\)", pv).group(1), 16)
    pgen = int(re.search(r"const MULTIPLICATIVE_GENERATOR: Self = Self\((0x[0-9a-f]+)\)", pv).group(1), 16)
    mont = int(re.search(r"self \* Self\((0x[0-9a-f]+)\)", pv).group(1), 16)

    out = {
        "source": "IrreducibleOSS/binius @ 47675e1 (see module docstring for file:line)",
        "mul_kats": kats,
        "generators": gens,
        "aes_generators": aes_gens,
        "aes_to_binary": aes_to_bin,
        "binary_to_aes": bin_to_aes,
        "binary_to_polyval": [hex(x) for x in b2p],
        "polyval_to_binary": [hex(x) for x in p2b],
        "polyval_one": hex(one),
        "polyval_generator": hex(pgen),
        "polyval_to_montgomery_const": hex(mont),
        "polyval_mul_kat": [hex(x) for x in mul_kat],
        "polyval_sqr_kat": [hex(x) for x in sqr_kat],
    }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, {k: len(v) for k, v in kats.items()})


if __name__ == "__main__":
    main()
