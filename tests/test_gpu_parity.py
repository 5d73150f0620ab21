"""GPU parity: the B200 engine against the CPU oracle, through the C ABI.

Bit-exactness bar: the engine walks every floating-point sum in the reference's order
(see k_lpc_model.cuh), so whole FILE IMAGES must be byte-identical to the oracle's -- not
only the integer stages.  (north_star would allow 0.1 % size drift for the FP stage; we do
not use that allowance.)"""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _b200():
    import b200flac
    return b200flac


def _encode_b200(tmp_path, pcm, rate, ch, bps, opts, name="o.flac"):
    b = _b200()
    kw = {k: v for k, v in opts.items() if k != "padding_size"}
    p = b.make_params(rate, ch, bps, **kw)
    path = os.path.join(str(tmp_path), name)
    n = len(pcm) // (ch * (bps // 8))
    b.encode_file(path, p, pcm, n, padding_size=opts["padding_size"])
    with open(path, "rb") as fh:
        return fh.read()


def _first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    idx = np.nonzero(x)[0]
    return int(idx[0]) if len(idx) else n


def _check(tmp_path, pcm, rate, ch, bps, opts):
    want = helpers.oracle_encode(pcm, rate, ch, bps, opts)
    got = _encode_b200(tmp_path, pcm, rate, ch, bps, opts)
    assert len(got) == len(want) and got == want, \
        "file differs at byte %d (lengths %d vs %d) for %r" % (_first_diff(got, want), len(got), len(want), opts)


GRID = [
    # (rate, channels, bps, frames, options)
    (44100, 2, 16, 4096 * 6 + 100, dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=3)),
    (44100, 2, 16, 4096 * 6 + 100, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6,
                                       adaptive_mid_side=True)),
    (44100, 2, 16, 4096 * 4 + 7, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6,
                                     mid_side=True, exhaustive_model_search=True)),
    (44100, 2, 16, 1152 * 9 + 3, dict(block_size=1152, max_lpc_order=0, max_residual_partition_order=3)),
    (44100, 2, 16, 1152 * 9 + 3, dict(block_size=1152, max_lpc_order=0, max_residual_partition_order=3,
                                     adaptive_mid_side=True)),
    (96000, 2, 24, 4096 * 4 + 99, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8,
                                      mid_side=True, exhaustive_model_search=True)),
    (96000, 6, 24, 4608 * 3 + 500, dict(block_size=4608, max_lpc_order=12, max_residual_partition_order=6)),
    (48000, 1, 8, 5000, dict(block_size=256, max_lpc_order=6, max_residual_partition_order=4)),
    (44100, 2, 16, 20000, dict(block_size=4096, max_lpc_order=32, max_residual_partition_order=8, mid_side=True)),
    (44100, 3, 16, 9000, dict(block_size=2304, max_lpc_order=16, max_residual_partition_order=5)),
]


@pytest.mark.parametrize("case", range(len(GRID)))
def test_synth_file_identical(case, tmp_path, built):
    rate, ch, bps, n, o = GRID[case]
    pcm = helpers.synth_pcm(1234 + case, ch, bps, n)
    _check(tmp_path, pcm, rate, ch, bps, helpers.options(**o))


def test_device_synth_matches_oracle_generator(built):
    b = _b200()
    import ctypes as C
    for ch, bps, n in ((2, 16, 10000), (6, 24, 5000), (1, 8, 3000)):
        nbytes = n * ch * (bps // 8)
        d = b.lib().b200flac_device_alloc(0, nbytes)
        assert d
        assert b.lib().b200flac_device_synth_pcm(0, d, 99, ch, bps, 12345, n) == 0
        host = np.empty(nbytes, dtype=np.uint8)
        assert b.lib().b200flac_device_download(0, host.ctypes.data, d, nbytes) == 0
        b.lib().b200flac_device_free(0, d)
        assert host.tobytes() == helpers.synth_pcm(99, ch, bps, n, first_frame=12345)
