#!/usr/bin/env python
"""Generates tests/golden/tta_golden.json from the COMPILED REFERENCE TTA encoder (oracle/_ref/ttaenc, built by
`make -C oracle ref` from the unmodified /root/reference/src/encoders/tta.c): sha256 and length of its output
file for deterministic inputs; the reference's own decoder (oracle/_ref/ttadec) must return the input.

    python tests/golden/make_tta_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from tta_cases import TTA_CASES, tta_case_pcm  # noqa: E402


def main():
    assert helpers.have_tta_ref(), "build oracle/_ref first: make -C oracle ref"
    out = []
    for case in TTA_CASES:
        pcm = tta_case_pcm(case)
        data = helpers.ref_tta_encode(pcm, case["rate"], case["channels"], case["bps"])
        assert helpers.ref_tta_decode(data) == pcm, case["name"]
        out.append({"name": case["name"], "sha256": hashlib.sha256(data).hexdigest(), "length": len(data),
                    "pcm_sha256": hashlib.sha256(pcm).hexdigest()})
        print("%-32s %9d bytes" % (case["name"], len(data)))
    with open(os.path.join(HERE, "tta_golden.json"), "w") as fh:
        json.dump({"generator": "oracle/_ref/ttaenc (reference src/encoders/tta.c, -O2 -DNDEBUG -DSTANDALONE)", "cases": out},
                  fh, indent=1)


if __name__ == "__main__":
    main()
