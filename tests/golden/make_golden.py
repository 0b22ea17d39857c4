#!/usr/bin/env python3
"""Extract the reference's own golden vectors for BLS12-381 Fr into a JSON fixture.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Sources (all under /root/reference):
* plugins/arkworks/src/poseidon/parameters_hardcoded_test/lfsr_values   (asserted at
  plugins/arkworks/src/poseidon/test.rs:59-65) -- 189 Grain-LFSR round constants for
  (255 bits, width 3, 8 full rounds, 55 partial rounds)
* plugins/arkworks/src/poseidon/mds_hardcoded_tests/width2..width12      (asserted at
  plugins/arkworks/src/poseidon/test.rs:477-497) -- Cauchy MDS matrices
* plugins/arkworks/src/poseidon/permutation_hardcoded_test/width3 and
  openzl-tutorials/src/poseidon.rs:388-401 -- width-3 permutation KAT on input [3, 1, 2]
Only decimal constants are extracted; no reference source code is copied.
"""
import json
import os
import re

REF = "/root/reference"
POS = os.path.join(REF, "plugins/arkworks/src/poseidon")
NUM = re.compile(r'field_new!\(\s*Fr,\s*"(\d+)"\s*\)')


def numbers(path):
    with open(path) as fh:
        return [int(m) for m in NUM.findall(fh.read())]


def main():
    out = {
        "field": "bls12_381_fr",
        "lfsr_values": {"modulus_bits": 255, "width": 3, "full_rounds": 8, "partial_rounds": 55,
                        "values": [str(v) for v in numbers(os.path.join(POS, "parameters_hardcoded_test/lfsr_values"))]},
        "mds": {},
        "permutation_width3": {"input": ["3", "1", "2"],
                               "output": [str(v) for v in numbers(os.path.join(POS, "permutation_hardcoded_test/width3"))]},
    }
    for t in range(2, 13):
        vals = numbers(os.path.join(POS, f"mds_hardcoded_tests/width{t}"))
        assert len(vals) == t * t, (t, len(vals))
        out["mds"][str(t)] = [str(v) for v in vals]
    tut = numbers(os.path.join(REF, "openzl-tutorials/src/poseidon.rs"))
    out["permutation_width3"]["tutorial_output"] = [str(v) for v in tut[-3:]]
    assert len(out["lfsr_values"]["values"]) == 189
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poseidon_bls12_381_fr.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", dst)


if __name__ == "__main__":
    main()
