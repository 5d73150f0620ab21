#!/usr/bin/env python
"""Generates tests/golden/alac_golden.json from the COMPILED REFERENCE ALAC encoder (oracle/_ref/alacenc, built by
`make -C oracle ref` from the unmodified /root/reference/src/encoders/alac.c): sha256 and length of the mdat atom it
writes for deterministic inputs (default options: history 10 / 40, maximum k 14, leftweights 0..4).

    python tests/golden/make_alac_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from alac_cases import ALAC_CASES, alac_case_pcm  # noqa: E402


def main():
    assert helpers.have_alac_ref(), "build oracle/_ref first: make -C oracle ref"
    out = []
    for case in ALAC_CASES:
        pcm = alac_case_pcm(case)
        data = helpers.ref_alac_encode(pcm, case["channels"], case["bps"], case["block_size"])
        out.append({"name": case["name"], "sha256": hashlib.sha256(data).hexdigest(), "length": len(data),
                    "pcm_sha256": hashlib.sha256(pcm).hexdigest()})
        print("%-28s %9d bytes" % (case["name"], len(data)))
    with open(os.path.join(HERE, "alac_golden.json"), "w") as fh:
        json.dump({"generator": "oracle/_ref/alacenc (reference src/encoders/alac.c, -O2 -DNDEBUG -DSTANDALONE)", "cases": out},
                  fh, indent=1)


if __name__ == "__main__":
    main()
