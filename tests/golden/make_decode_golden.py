#!/usr/bin/env python
"""Builds tests/golden/decode_golden.json and tests/golden/flac/: the reference's own FLAC fixtures
(/root/reference/test/*.flac -- files written by other encoders: libFLAC tones, all subframe types, metadata
in unusual order, ID3-prefixed, blank MD5) copied as binary fixtures, and what the COMPILED REFERENCE DECODER
(oracle/_ref/flacdec, src/decoders/flac.c built as is) does with each: exit status, stderr, length and
sha256 of the PCM.  Run in the build container, where /root/reference exists; the GPU tests only read the
committed results."""
import glob
import hashlib
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "flacdec")
SKIP = {"tone.flac"}  # 440 KB; tone1..8 cover the same encoder


def main():
    out = {}
    for src in sorted(glob.glob("/root/reference/test/*.flac")):
        name = os.path.basename(src)
        if name in SKIP:
            continue
        shutil.copyfile(src, os.path.join(HERE, "flac", name))
        r = subprocess.run([REF, src], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        out[name] = {"rc": r.returncode, "stderr": r.stderr.decode(), "pcm_bytes": len(r.stdout),
                     "pcm_sha256": hashlib.sha256(r.stdout).hexdigest()}
    with open(os.path.join(HERE, "decode_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("%d fixtures" % len(out))


if __name__ == "__main__":
    main()
