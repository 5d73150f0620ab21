#!/usr/bin/env python
"""Generates tests/golden/golden.json from the COMPILED REFERENCE (oracle/_ref/flacenc, built by
`make -C oracle ref` from the unmodified sources under /root/reference/src).

Each case names a deterministic input (the integer synthetic generator or a restated reference
test stream) and the reference encoder's options; the manifest records the sha256, length and the
first 64 bytes after the metadata of the reference's output.  /root/reference is only needed to
build oracle/_ref, i.e. to REGENERATE this file; the tests read the committed manifest.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from golden_cases import CASES, case_pcm  # noqa: E402


def main():
    assert helpers.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = []
    for case in CASES:
        pcm = case_pcm(case)
        opts = helpers.options(**case["options"])
        flac = helpers.ref_encode(pcm, case["rate"], case["channels"], case["bps"], opts)
        assert helpers.ref_decode(flac) == pcm, case["name"]
        ff = helpers.first_frame_offset(flac)
        out.append({"name": case["name"], "sha256": hashlib.sha256(flac).hexdigest(), "length": len(flac),
                    "pcm_sha256": hashlib.sha256(pcm).hexdigest(),
                    "first_frame_bytes": flac[ff:ff + 64].hex()})
        print("%-40s %8d bytes" % (case["name"], len(flac)))
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump({"generator": "oracle/_ref/flacenc (reference src/encoders/flac.c, -O2 -DNDEBUG -DSTANDALONE, "
                                "VERSION=2.22alpha1)", "cases": out}, fh, indent=1)


if __name__ == "__main__":
    main()
