"""Extracts the literal golden vectors that are too long to retype from the
reference's own test files into JSON fixtures.  Run in the build container (where
/root/reference exists); the emitted JSON is committed, this script documents how.

  python tests/golden/make_golden.py
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def unpackbits_test1():
    """test/indexing/codecs/residual.jl:277-816 -- `_unpackbits` Test 1: 64 packed
    bytes (nbits = 1) and the expected 512 unpacked bits."""
    src = open(os.path.join(REF, "test/indexing/codecs/residual.jl")).read().split("\n")
    # locate the testset
    start = next(i for i, l in enumerate(src) if '@testset "_unpackbits" begin' in l)
    end = next(i for i in range(start, len(src)) if "unpacked_bits = _unpackbits(packed_bits, nbits)" in src[i])
    block = "\n".join(src[start:end])
    packed_part, expected_part = block.split("expected = reshape(", 1)
    packed = [int(b, 2) for b in re.findall(r"0b([01]{8})", packed_part)]
    bits_txt = expected_part.split("Bool[", 1)[1].split("]", 1)[0]
    bits = [int(x) for x in re.findall(r"[01]", bits_txt)]
    assert len(packed) == 64 and len(bits) == 512, (len(packed), len(bits))
    return {"source": "test/indexing/codecs/residual.jl:277-816", "nbits": 1,
            "packed_shape_julia": [1, 64], "packed": packed, "expected_bits": bits}


if __name__ == "__main__":
    with open(os.path.join(OUT, "unpackbits_test1.json"), "w") as f:
        json.dump(unpackbits_test1(), f)
    print("wrote unpackbits_test1.json")
