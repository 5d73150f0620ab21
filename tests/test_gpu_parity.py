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


@pytest.mark.parametrize("bs", [576, 1024])
@pytest.mark.parametrize("chunk,lookahead", [(7, 1), (16, 3), (33, 2), (0, 3)])
def test_chunk_pipeline_same_frames(chunk, lookahead, bs, built):
    """b200flac_encoder_set_chunking: a batch run as a pipeline of chunks over three streams (model kernels
    ahead on a high-priority stream, offsets chained through the running totals) yields the oracle's frames,
    for chunk sizes that do and do not divide the batch, with short blocks (two segment tails) inside;
    block 576 runs on k_analyze_v2, block 1024 on k_analyze_v3 (+ v2 for the tails)"""
    b = _b200()
    o = helpers.options(block_size=bs, max_lpc_order=12, max_residual_partition_order=4, adaptive_mid_side=True)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    n1, n2 = bs * 83 + 100, bs * 60 + 7
    pcm = helpers.synth_pcm(321, 2, 16, n1 + n2)
    enc = b.Encoder(p, max_pcm_frames_per_batch=n1 + n2, n_slots=1)
    enc.set_chunking(chunk, lookahead)
    for _ in range(2):     # twice: the pipeline's streams and events are reused
        out, fbytes, fpcm = enc.encode(pcm, n1 + n2, segments=[(0, n1, 0), (n1, n2, 500)])
    w1, _ = helpers.oracle_encode_range(pcm[:n1 * 4], 44100, 2, 16, o, 0)
    w2, _ = helpers.oracle_encode_range(pcm[n1 * 4:], 44100, 2, 16, o, 500)
    assert out.tobytes() == w1 + w2
    assert int(fbytes.sum()) == len(w1) + len(w2) and list(fpcm) == [bs] * 83 + [100] + [bs] * 60 + [7]
    assert (len(enc.kernel_ms(0)) == 0) == (chunk != 0)
    enc.close()


def test_encode_file_from_pinned_memory_in_place(tmp_path, built):
    """b200flac_encode_file with the PCM in page-locked memory (b200flac_host_alloc): batches go to the device and
    to the MD5 thread straight from the caller's buffer -- same file as from pageable memory and as the oracle's;
    several batches (block 256 -> 2048-block batches) and a ragged tail"""
    import ctypes as C
    b = _b200()
    L = b.lib()
    o = helpers.options(block_size=256, max_lpc_order=4, max_residual_partition_order=3, adaptive_mid_side=True)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    n = 256 * 2048 * 3 + 256 * 11 + 5
    pcm = helpers.synth_pcm(90, 2, 16, n)
    h = L.b200flac_host_alloc(len(pcm))
    assert h
    C.memmove(h, pcm, len(pcm))
    path = os.path.join(str(tmp_path), "pinned.flac")
    dev = (C.c_int * 1)(0)
    assert L.b200flac_encode_file(os.fsencode(path), C.byref(p), 4096, None, h, n, dev, 1) == 0
    L.b200flac_host_free(h)
    assert open(path, "rb").read() == helpers.oracle_encode(pcm, 44100, 2, 16, o)


def test_device_resident_async_two_slots(built):
    """b200flac_encoder_submit_device / collect_device: two batches in flight on two slots give what the
    synchronous call gives, and the oracle's frames"""
    b = _b200()
    L = b.lib()
    o = helpers.options(block_size=1024, max_lpc_order=8, max_residual_partition_order=4, adaptive_mid_side=True)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    n = 1024 * 50 + 300
    enc = b.Encoder(p, max_pcm_frames_per_batch=n, n_slots=2)
    cap = enc.output_bound(n, 1)
    bufs = []
    for seed in (5, 6):
        d_pcm, d_out = L.b200flac_device_alloc(0, n * 4), L.b200flac_device_alloc(0, cap)
        assert L.b200flac_device_synth_pcm(0, d_pcm, seed, 2, 16, 0, n) == 0
        bufs.append((seed, d_pcm, d_out))
    for rep in range(2):
        for slot, (seed, d_pcm, d_out) in enumerate(bufs):
            enc.submit_device(d_pcm, [(0, n, 0)], d_out, cap, slot=slot)
        with pytest.raises(b.B200FlacError, match="busy"):
            enc.submit_device(bufs[0][1], [(0, n, 0)], bufs[0][2], cap, slot=0)
        for slot, (seed, d_pcm, d_out) in enumerate(bufs):
            nbytes, nfr, ms = enc.collect_device(slot=slot)
            host = np.empty(nbytes, dtype=np.uint8)
            assert L.b200flac_device_download(0, host.ctypes.data, d_out, nbytes) == 0
            want, _ = helpers.oracle_encode_range(helpers.synth_pcm(seed, 2, 16, n), 44100, 2, 16, o, 0)
            assert nfr == 51 and host.tobytes() == want and ms > 0
    with pytest.raises(b.B200FlacError, match="no batch"):
        enc.collect_device(slot=0)
    for _, d_pcm, d_out in bufs:
        L.b200flac_device_free(0, d_pcm)
        L.b200flac_device_free(0, d_out)
    enc.close()


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


# ---------------------------------------------------------------------------------------------
# golden vectors: outputs of the COMPILED REFERENCE (tests/golden/golden.json)
# ---------------------------------------------------------------------------------------------
import hashlib  # noqa: E402
import json  # noqa: E402

from golden.golden_cases import CASES, LEVELS, case_pcm  # noqa: E402

with open(os.path.join(helpers.GOLDEN, "golden.json")) as _fh:
    GOLD = {c["name"]: c for c in json.load(_fh)["cases"]}


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_reference_outputs(case, tmp_path, built):
    """the engine's file == the reference encoder's file (sha256 recorded from oracle/_ref/flacenc)"""
    g = GOLD[case["name"]]
    pcm = case_pcm(case)
    assert hashlib.sha256(pcm).hexdigest() == g["pcm_sha256"], "input drifted"
    got = _encode_b200(tmp_path, pcm, case["rate"], case["channels"], case["bps"], helpers.options(**case["options"]))
    ff = helpers.first_frame_offset(got)
    assert got[ff:ff + 64].hex() == g["first_frame_bytes"]
    assert len(got) == g["length"]
    assert hashlib.sha256(got).hexdigest() == g["sha256"]


@pytest.mark.parametrize("mask", range(16))
def test_disable_subframe_flags(mask, tmp_path, built):
    """the reference's debug switches (flac.c:732-809 truth table); the oracle restates them"""
    o = helpers.options(block_size=1024, max_lpc_order=8, max_residual_partition_order=4, mid_side=True,
                        disable_verbatim_subframes=bool(mask & 1), disable_constant_subframes=bool(mask & 2),
                        disable_fixed_subframes=bool(mask & 4), disable_lpc_subframes=bool(mask & 8))
    # noise (VERBATIM wins), a constant block, silence and a tone, so every branch is reachable
    rng = np.random.RandomState(mask)
    parts = [rng.randint(-32768, 32768, size=2048), np.full(2048, 1234), np.zeros(2048, dtype=np.int64),
             (8000 * np.sin(np.arange(4096) * 0.05)).astype(np.int64)]
    pcm = helpers.pack_pcm(np.concatenate(parts).astype(np.int32), 16)
    _check(tmp_path, pcm, 44100, 2, 16, o)


def test_frame_layer_many_segments(built):
    """b200flac_encoder_encode with several streams in one batch (config #5 shape): every segment is
    blocked on its own and numbered from its own first_frame_number"""
    b = _b200()
    o = helpers.options(block_size=1152, max_lpc_order=8, max_residual_partition_order=4, adaptive_mid_side=True)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    lens = [1152 * 3 + 7, 1152 * 2, 5, 1152 * 4 + 1151]
    firsts = [0, 100, 127, 70000]
    tracks = [helpers.synth_pcm(900 + i, 2, 16, n) for i, n in enumerate(lens)]
    pcm = b"".join(tracks)
    segs, pos = [], 0
    for n, f in zip(lens, firsts):
        segs.append((pos, n, f))
        pos += n
    enc = b.Encoder(p, max_pcm_frames_per_batch=pos, n_slots=1)
    out, fbytes, fpcm = enc.encode(pcm, pos, segments=segs)
    want, want_sizes = b"", []
    for t, f in zip(tracks, firsts):
        fr, sz = helpers.oracle_encode_range(t, 44100, 2, 16, o, f)
        want += fr
        want_sizes += sz
    assert fbytes.tolist() == want_sizes
    assert out.tobytes() == want
    assert fpcm.tolist() == [1152, 1152, 1152, 7, 1152, 1152, 5, 1152, 1152, 1152, 1152, 1151]
    enc.close()


def test_stream_write_chunking_and_batches(tmp_path, built):
    """stream layer: odd write sizes, several device batches in flight (block 256 -> 2048-block batches)"""
    b = _b200()
    o = helpers.options(block_size=256, max_lpc_order=4, max_residual_partition_order=3, adaptive_mid_side=True)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    n = 256 * 2048 * 2 + 256 * 700 + 33
    pcm = helpers.synth_pcm(77, 2, 16, n)
    path = os.path.join(str(tmp_path), "s.flac")
    s = b.Stream(path, p)
    pos, k = 0, 0
    sizes = [1, 4099, 100000, 17, 256 * 2048 * 4]
    while pos < len(pcm):
        take = min(sizes[k % len(sizes)] * 4, len(pcm) - pos)
        s.write(pcm[pos:pos + take])
        pos += take
        k += 1
    offs = s.close()
    want, want_offs = helpers.oracle_encode(pcm, 44100, 2, 16, o, want_offsets=True)
    got = open(path, "rb").read()
    assert got == want
    assert offs == want_offs


def test_stream_end_block_short_reads(tmp_path, built):
    """a reader that returns short blocks mid-stream: the reference encodes each read as its own frame
    (flac.c:247,525)"""
    b = _b200()
    o = helpers.options(block_size=512, max_lpc_order=6, max_residual_partition_order=3)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(48000, 1, 16, **kw)
    reads = [512, 512, 100, 512, 7, 512, 512, 300]
    pcm = helpers.synth_pcm(5, 1, 16, sum(reads))
    path = os.path.join(str(tmp_path), "e.flac")
    s = b.Stream(path, p)
    pos = 0
    for r in reads:
        s.write(pcm[pos * 2:(pos + r) * 2])
        if r < 512:
            s.end_block()
        pos += r
    offs = s.close()
    assert [n for _, n in offs] == reads
    # frames must equal the oracle's, read by read, with consecutive frame numbers
    want, pos = b"", 0
    for i, r in enumerate(reads):
        fr, _ = helpers.oracle_encode_range(pcm[pos * 2:(pos + r) * 2], 48000, 1, 16, o, i)
        want += fr
        pos += r
    got = open(path, "rb").read()
    assert got[helpers.first_frame_offset(got):] == want
    if helpers.have_ref():
        assert helpers.ref_decode(got) == pcm


def test_stream_every_read_short(tmp_path, built):
    """a reader whose every read is shorter than block_size (a non-buffered pcmreader): thousands of
    one-frame segments, more than one encoder batch holds tails for -- the stream layer must end
    batches early instead of failing (round-1 advisor finding; reference behaviour flac.c:247,525)"""
    b = _b200()
    o = helpers.options(block_size=4096, max_lpc_order=4, max_residual_partition_order=2)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 1, 16, **kw)
    n_reads, r = 2600, 96
    pcm = helpers.synth_pcm(9, 1, 16, n_reads * r)
    path = os.path.join(str(tmp_path), "short.flac")
    s = b.Stream(path, p)
    for i in range(n_reads):
        s.write(pcm[i * r * 2:(i + 1) * r * 2])
        s.end_block()
    offs = s.close()
    assert [n for _, n in offs] == [r] * n_reads
    got = open(path, "rb").read()
    frames = got[helpers.first_frame_offset(got):]
    # spot-check frames against the oracle (frame numbers cross the 1-, 2- and 3-byte UTF-8 forms)
    for i in (0, 1, 127, 128, 1023, 1024, 2047, 2048, n_reads - 1):
        fr, _ = helpers.oracle_encode_range(pcm[i * r * 2:(i + 1) * r * 2], 44100, 1, 16, o, i)
        a = offs[i][0]
        assert frames[a:a + len(fr)] == fr, i
    if helpers.have_ref():
        assert helpers.ref_decode(got) == pcm


def test_stream_frame_ranges_over_two_devices(tmp_path, built):
    """ONE stream whose frame ranges go to several GPUs (b200flac_stream_open(devices, n)): batches of
    consecutive blocks are handed to the devices in turn, their frames concatenated in order on the host, frame
    numbers and the offsets list continuous, one MD5 over the whole stream.  Needs two devices."""
    b = _b200()
    if b.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    # (a) small blocks, 5+ batches of 2048 blocks: bytes equal to the oracle's
    o = helpers.options(block_size=256, max_lpc_order=4, max_residual_partition_order=3, adaptive_mid_side=True)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    n = 256 * 2048 * 5 + 256 * 300 + 77
    pcm = helpers.synth_pcm(88, 2, 16, n)
    path = os.path.join(str(tmp_path), "two.flac")
    s = b.Stream(path, p, devices=[0, 1])
    s.write(pcm)
    offs = s.close()
    want, want_offs = helpers.oracle_encode(pcm, 44100, 2, 16, o, want_offsets=True)
    assert open(path, "rb").read() == want
    assert offs == want_offs
    # (b) config 4's shape (96 kHz/24-bit 5.1, block 4608), 2100 blocks: the two-device file equals the one-device
    # file, and the reference decoder returns the PCM
    o4 = helpers.options(block_size=4608, max_lpc_order=12, max_residual_partition_order=6)
    kw4 = {k: v for k, v in o4.items() if k != "padding_size"}
    p4 = b.make_params(96000, 6, 24, **kw4)
    n4 = 4608 * 2100 + 1000
    pcm4 = b.synth_pcm(89, 6, 24, n4)
    one, two = os.path.join(str(tmp_path), "c4_one.flac"), os.path.join(str(tmp_path), "c4_two.flac")
    b.encode_file(one, p4, pcm4, n4, devices=[0])
    b.encode_file(two, p4, pcm4, n4, devices=[1, 0])
    a, c = open(one, "rb").read(), open(two, "rb").read()
    assert a == c
    info, back = b.decode(c)
    assert back == pcm4 and info.total_pcm_frames == n4


def test_lpc_order_estimate_stress(tmp_path, built):
    """The order estimate (flac.c:1233-1268) is the one place where the device's libm (log) is not the host's: an
    ulp of difference could only matter where two orders' estimates agree to ~1e-15.  6,000 short blocks of
    synthetic resonances whose LPC error curves are deliberately flat (pole radius close to 1, orders beyond the
    model all equally good), so that neighbouring orders' estimates are as close as real signals make them: every
    block's chosen order, coefficients and bytes must equal the oracle's (which runs the host libm)."""
    from scipy.signal import lfilter
    rng = np.random.RandomState(2024)
    n_blocks, bs = 6000, 192
    x = np.zeros(n_blocks * bs)
    e = rng.standard_normal(n_blocks * bs)
    for b in range(n_blocks):
        # an AR(2) or AR(4) resonance per block, excitation level varied over 60 dB
        r1, f1 = rng.uniform(0.90, 0.9995), rng.uniform(0.01, 0.45)
        a = [2 * r1 * np.cos(2 * np.pi * f1), -r1 * r1]
        if b & 1:
            r2, f2 = rng.uniform(0.5, 0.999), rng.uniform(0.01, 0.45)
            p = np.polymul([1, -a[0], -a[1]], [1, -2 * r2 * np.cos(2 * np.pi * f2), r2 * r2])
            a = list(-p[1:])
        g = 10 ** rng.uniform(0, 3)
        y = lfilter([g], np.concatenate([[1.0], -np.asarray(a)]), e[b * bs:(b + 1) * bs])
        m = np.max(np.abs(y))
        x[b * bs:(b + 1) * bs] = y * (rng.uniform(200, 30000) / m if m > 0 else 0)
    pcm = helpers.pack_pcm(np.round(x).astype(np.int32), 16)
    for o in (dict(block_size=bs, max_lpc_order=12, max_residual_partition_order=3),
              dict(block_size=bs, max_lpc_order=32, max_residual_partition_order=2)):
        _check(tmp_path, pcm, 44100, 1, 16, helpers.options(**o))


def test_error_paths(tmp_path, built):
    b = _b200()
    p = b.make_params()
    with pytest.raises(b.B200FlacError):
        b.encode_file(os.path.join(str(tmp_path), "no", "such", "dir", "x.flac"), p, b"\0" * 16, 4)
    bad = b.make_params(bits_per_sample=12)
    with pytest.raises(b.B200FlacError):
        b.Encoder(bad)
    with pytest.raises(b.B200FlacError):
        b.Encoder(b.make_params(channels=9))
    # empty stream: head only, like the reference
    path = os.path.join(str(tmp_path), "empty.flac")
    b.encode_file(path, p, b"", 0)
    assert open(path, "rb").read() == helpers.oracle_encode(b"", 44100, 2, 16, helpers.options(
        block_size=4096, max_lpc_order=8, max_residual_partition_order=5))


def test_full_size_properties(tmp_path, built):
    """size-independent properties at a realistic size (10 min of stereo): the reference's own
    decoder accepts the stream (CRC-16 of every frame, STREAMINFO MD5) and returns the input;
    frame sizes sum to the file; where the compiled reference is present the bytes are equal too"""
    b = _b200()
    n = 44100 * 600
    o = helpers.options(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)
    pcm = helpers.synth_pcm(1235, 2, 16, n)
    got = _encode_b200(tmp_path, pcm, 44100, 2, 16, o)
    si = helpers.streaminfo(got)
    assert si["total_samples"] == n and si["md5"] == hashlib.md5(pcm).digest()
    if helpers.have_ref():
        assert helpers.ref_decode(got) == pcm
        assert got == helpers.ref_encode(pcm, 44100, 2, 16, o)


# ---------------------------------------------------------------------------------------------
# kernel generations: k_analyze_v3 / k_pack_v3 (full-length blocks of the common shapes) against the
# oracle on shapes that reach their special cases, and against k_analyze_v2 / k_pack_v2 / the generic
# kernels, which the same library selects through tuning knobs read at encoder creation
# ---------------------------------------------------------------------------------------------
V3_GRID = [
    # 8192-sample blocks: 256-thread CTAs, partition order 7 (finest partition = 2 thread runs)
    (44100, 2, 16, 8192 * 3 + 4000, dict(block_size=8192, max_lpc_order=12, max_residual_partition_order=7, mid_side=True)),
    # 4608: 24 samples per thread, a finest partition is three thread runs
    (48000, 2, 16, 4608 * 4 + 11, dict(block_size=4608, max_lpc_order=10, max_residual_partition_order=6, adaptive_mid_side=True)),
    # one warp per unit, four partitions
    (44100, 1, 16, 512 * 9 + 1, dict(block_size=512, max_lpc_order=4, max_residual_partition_order=2)),
    # 24-bit: 64-bit FIXED sums, 64-bit LPC accumulation, Rice parameters above 14 (coding method 1)
    (96000, 2, 24, 4096 * 3 + 9, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True)),
    (96000, 6, 24, 4096 * 2 + 500, dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=5)),
    # orders above 12 (32-tap kernels), 16 (zero padded to 32)
    (44100, 2, 16, 4096 * 3, dict(block_size=4096, max_lpc_order=32, max_residual_partition_order=6, mid_side=True)),
    (44100, 2, 16, 4096 * 3, dict(block_size=4096, max_lpc_order=16, max_residual_partition_order=4)),
    # exhaustive order search on the v3 kernel, orders above 12 (32-tap residual), 8192-sample blocks
    (44100, 2, 16, 4096 * 2 + 50, dict(block_size=4096, max_lpc_order=16, max_residual_partition_order=6, mid_side=True,
                                      exhaustive_model_search=True)),
    (44100, 2, 16, 8192 * 2 + 50, dict(block_size=8192, max_lpc_order=8, max_residual_partition_order=6, adaptive_mid_side=True,
                                      exhaustive_model_search=True)),
    (96000, 2, 24, 4096 * 2 + 50, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True,
                                      exhaustive_model_search=True)),
    # partition order 0 only
    (44100, 2, 16, 4096 * 2 + 77, dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=0, adaptive_mid_side=True)),
    # partition order 8 (a finest partition is half a thread run: the SUB = 2 instantiations), with and without
    # the exhaustive search, 16 and 24 bits, BASELINE config 3's options
    (44100, 2, 16, 4096 * 4 + 321, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True)),
    (44100, 2, 16, 4096 * 4 + 321, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True,
                                       exhaustive_model_search=True)),
    (96000, 2, 24, 4096 * 4 + 99, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True,
                                      exhaustive_model_search=True)),
    (96000, 1, 24, 4096 * 3 + 5, dict(block_size=4096, max_lpc_order=16, max_residual_partition_order=8)),
    (48000, 6, 16, 4096 * 2 + 1000, dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=8,
                                        exhaustive_model_search=True)),
]


@pytest.mark.parametrize("case", range(len(V3_GRID)))
def test_v3_shapes_identical(case, tmp_path, built):
    rate, ch, bps, n, o = V3_GRID[case]
    pcm = helpers.synth_pcm(4321 + case, ch, bps, n)
    _check(tmp_path, pcm, rate, ch, bps, helpers.options(**o))


def _special_signal(bps, n_blocks, block):
    """blocks that reach the rare branches: wasted bits, constant, silence, full scale noise (VERBATIM),
    a near-constant block (Rice parameter 0), alternating extremes (largest FIXED residuals)"""
    rng = np.random.RandomState(7)
    full = 1 << (bps - 1)
    t = np.arange(block)
    parts = [
        ((3000 * np.sin(t * 0.03)).astype(np.int64) << 3),                  # 3 wasted bits
        np.full(block, -1234, dtype=np.int64),                              # constant
        np.zeros(block, dtype=np.int64),                                    # silence (constant 0)
        rng.randint(-full, full, size=block).astype(np.int64),              # noise: VERBATIM
        (t % 7 == 0).astype(np.int64),                                      # k = 0 partitions
        np.where(t & 1, full - 1, -full).astype(np.int64),                  # alternating extremes
        (rng.randint(-full // 4, full // 4, size=block).astype(np.int64) & ~1),  # noise with 1 wasted bit
        (20000 * np.sin(t * 0.2) * np.exp(-t / 900.0)).astype(np.int64),    # decaying tone: k varies over partitions
    ]
    mono = np.concatenate([parts[i % len(parts)] for i in range(n_blocks)])
    return mono


@pytest.mark.parametrize("bps,opts", [
    (16, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)),
    (16, dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=5, mid_side=True)),
    (24, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True)),
    (16, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True)),
    (16, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, adaptive_mid_side=True, exhaustive_model_search=True)),
    (24, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True, exhaustive_model_search=True)),
])
def test_v3_special_blocks(bps, opts, tmp_path, built):
    block = opts["block_size"]
    left = _special_signal(bps, 10, block)
    right = np.roll(left, block * 3) // 2 + (np.arange(len(left)) % 5)      # a different block type per channel
    full = 1 << (bps - 1)
    right = np.clip(right, -full, full - 1)
    inter = np.empty(2 * len(left), dtype=np.int32)
    inter[0::2] = left
    inter[1::2] = right
    pcm = helpers.pack_pcm(inter, bps)
    _check(tmp_path, pcm, 44100, 2, bps, helpers.options(**opts))


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("case", [0, 1, 3, 4, 5])
def test_kernel_generations_agree(case, tmp_path, built):
    """the v3 kernels, the v2 kernels and the generic kernels write the same file"""
    rate, ch, bps, n, o = (GRID[1], V3_GRID[1], V3_GRID[3], V3_GRID[0], V3_GRID[12], V3_GRID[13])[case]
    pcm = helpers.synth_pcm(99 + case, ch, bps, n)
    opts = helpers.options(**o)
    a = _encode_b200(tmp_path, pcm, rate, ch, bps, opts, "a.flac")
    b = _with_env({"B200FLAC_NO_V3": "1", "B200FLAC_NO_P3": "1"},
                  lambda: _encode_b200(tmp_path, pcm, rate, ch, bps, opts, "b.flac"))
    c = _with_env({"B200FLAC_FORCE_GENERIC": "1"},
                  lambda: _encode_b200(tmp_path, pcm, rate, ch, bps, opts, "c.flac"))
    d = _with_env({"B200FLAC_LPC_G": "1"}, lambda: _encode_b200(tmp_path, pcm, rate, ch, bps, opts, "d.flac"))
    assert a == b, "v3 and v2 kernels differ at byte %d" % _first_diff(a, b)
    assert a == c, "v3 and generic kernels differ at byte %d" % _first_diff(a, c)
    assert a == d, "lag-split and unsplit autocorrelation differ at byte %d" % _first_diff(a, d)


def test_many_short_tracks_one_batch(built):
    """config #5 shape on the v3 kernels: tracks of 4096-sample blocks with short tails in one batch;
    tail frames go through k_analyze_v2, full ones through k_analyze_v3, all through k_pack_v3"""
    b = _b200()
    o = helpers.options(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    lens = [4096 * 2 + 100, 4096, 17, 4096 * 3 + 4095, 4096 * 1 + 1]
    tracks = [helpers.synth_pcm(300 + i, 2, 16, n) for i, n in enumerate(lens)]
    pcm = b"".join(tracks)
    segs, pos = [], 0
    for n in lens:
        segs.append((pos, n, 0))
        pos += n
    enc = b.Encoder(p, max_pcm_frames_per_batch=pos, n_slots=1)
    out, fbytes, fpcm = enc.encode(pcm, pos, segments=segs)
    want, want_sizes = b"", []
    for t in tracks:
        fr, sz = helpers.oracle_encode_range(t, 44100, 2, 16, o, 0)
        want += fr
        want_sizes += sz
    assert fbytes.tolist() == want_sizes
    assert out.tobytes() == want
    enc.close()


def test_standalone_driver_matches_reference_driver(tmp_path, built):
    """`b200flacenc [options] out.flac < pcm` writes the file `flacenc` (the compiled reference's driver,
    src/encoders/flac.c:1637-1804) writes for the same options"""
    import subprocess
    exe = os.path.join(helpers.ROOT, "python-audio-tools_b200", "b200flacenc")
    for ch, bps, rate, n, flags, o in (
        (2, 16, 44100, 4096 * 5 + 77, ["-M"], dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)),
        (1, 8, 8000, 3000, ["-B", "256", "-l", "4", "-R", "3"], dict(block_size=256, max_lpc_order=4, max_residual_partition_order=3)),
        (2, 24, 96000, 4096 * 2 + 5, ["-m", "-e", "-l", "8", "-R", "5"],
         dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=5, mid_side=True, exhaustive_model_search=True)),
    ):
        pcm = helpers.synth_pcm(70 + ch, ch, bps, n)
        out = os.path.join(str(tmp_path), "cli.flac")
        r = subprocess.run([exe, "-c", str(ch), "-b", str(bps), "-r", str(rate)] + flags + [out], input=pcm,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr
        assert b"Encoding from stdin using parameters:" in r.stdout
        got = open(out, "rb").read()
        want = helpers.ref_encode(pcm, rate, ch, bps, helpers.options(**o)) if helpers.have_ref() else \
            helpers.oracle_encode(pcm, rate, ch, bps, helpers.options(**o))
        assert got == want


# ---------------------------------------------------------------------------------------------
# many tracks -> many files in one call (b200flac_encode_files): batches of many segments, the
# STREAMINFO MD5 of every track computed on the device
# ---------------------------------------------------------------------------------------------
def test_device_md5_matches_hashlib(built):
    """the per-track MD5 kernel against hashlib at every padding boundary (0, 55|56, 63|64|65, 119|120, two blocks)
    and on a long string"""
    b = _b200()
    rng = np.random.RandomState(5)
    for n in (0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 121, 127, 128, 129, 1000, 4096 * 4 + 2, (1 << 20) + 3):
        data = rng.randint(0, 256, size=n).astype(np.uint8).tobytes()
        assert b.device_md5(data) == hashlib.md5(data).digest(), "length %d" % n


@pytest.mark.parametrize("shape", ["16bit_stereo", "24bit_6ch", "level8", "small_batches"])
def test_encode_files_equals_encode_file(shape, tmp_path, built):
    """every file of b200flac_encode_files is byte for byte what b200flac_encode_file writes for the same track:
    ragged lengths (empty track, shorter than a block, byte counts that are not a multiple of 64), several batches,
    region reuse in the device ring (B200FLAC_FILES_BATCH_MB=1, B200FLAC_FILES_RING=4)"""
    b = _b200()
    if shape == "24bit_6ch":
        rate, ch, bps = 96000, 6, 24
        o = helpers.options(block_size=4608, max_lpc_order=12, max_residual_partition_order=6)
        lengths = [4608 * 3 + 17, 100, 4608, 0, 4608 * 2 - 1, 9001]
    elif shape == "level8":
        rate, ch, bps = 44100, 2, 16
        o = helpers.options(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True,
                            exhaustive_model_search=True)
        lengths = [4096 * 5 + 333, 4096 * 2, 7, 4096 * 3 + 4095]
    else:
        rate, ch, bps = 44100, 2, 16
        o = helpers.options(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)
        lengths = [4096 * 4 + 99, 0, 1, 15, 16, 17, 4096, 4097, 4096 * 7 + 4095, 50000, 3, 4096 * 2 + 8]
        if shape == "small_batches":
            lengths = [30000 + 977 * i for i in range(40)]            # ~120-280 KB each: dozens of 1 MB batches
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(rate, ch, bps, **kw)
    pcms = [helpers.synth_pcm(700 + i, ch, bps, n) if n else b"" for i, n in enumerate(lengths)]
    want = []
    for i, (pcm, n) in enumerate(zip(pcms, lengths)):
        path = os.path.join(str(tmp_path), "one_%d.flac" % i)
        b.encode_file(path, p, pcm if n else b"\0", n, padding_size=o["padding_size"])
        want.append(open(path, "rb").read())
    names = [os.path.join(str(tmp_path), "many_%d.flac" % i) for i in range(len(lengths))]
    bufs = [np.frombuffer(pcm if n else b"\0" * 16, dtype=np.uint8).copy() for pcm, n in zip(pcms, lengths)]
    env = {"B200FLAC_FILES_BATCH_MB": "1", "B200FLAC_FILES_RING": "4"} if shape == "small_batches" else {}
    _with_env(env, lambda: b.encode_files(names, p, bufs, lengths, padding_size=o["padding_size"], device=0, host_threads=3))
    b.lib().b200flac_pool_clear()
    for i, name in enumerate(names):
        got = open(name, "rb").read()
        assert got == want[i], "track %d (%d PCM frames) differs at byte %d" % (i, lengths[i], _first_diff(got, want[i]))
    # and the hash in the STREAMINFO is the MD5 of the PCM (flac.c:187-188)
    for i, name in enumerate(names):
        assert open(name, "rb").read()[26:42] == hashlib.md5(pcms[i]).digest()


@pytest.mark.parametrize("mode", ["device_only", "host_only", "device_first"])
def test_encode_files_hash_sharing(mode, tmp_path, built):
    """the STREAMINFO MD5s are shared between the device (whole batches from the front of the list) and the pool's idle
    host threads (single tracks from its end): whoever hashes a track, the digest is hashlib's -- all on the device
    (B200FLAC_FILES_HOST_MD5=0), all on the host (a device rate so low that the whole list is the host's zone), and
    the two fronts meeting in the middle of the list (a host zone fixed at half of the job's PCM)"""
    b = _b200()
    o = helpers.options(block_size=4096, max_lpc_order=8, max_residual_partition_order=4)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    # (one track longer than a whole batch in the middle of the list, an empty one and two tiny ones at its end)
    lengths = [20000 + 1777 * (i % 9) for i in range(30)] + [800000] + [20000 + 1777 * (i % 9) for i in range(30)] + [0, 1, 4097]
    pcms = [helpers.synth_pcm(1300 + i, 2, 16, n) if n else b"" for i, n in enumerate(lengths)]
    bufs = [np.frombuffer(pcm if n else b"\0" * 16, dtype=np.uint8).copy() for pcm, n in zip(pcms, lengths)]
    names = [os.path.join(str(tmp_path), "h_%d.flac" % i) for i in range(len(lengths))]
    env = {"B200FLAC_FILES_BATCH_MB": "1", "B200FLAC_FILES_RING": "5"}
    env.update({"device_only": {"B200FLAC_FILES_HOST_MD5": "0"}, "host_only": {"B200FLAC_FILES_DEV_MD5_MBS": "0.000001"},
                "device_first": {"B200FLAC_FILES_HOST_ZONE_MB": "3"}}[mode])
    for threads in (1, 4):
        _with_env(env, lambda: b.encode_files(names, p, bufs, lengths, device=0, host_threads=threads))
        for i, name in enumerate(names):
            got = open(name, "rb").read()
            assert got[26:42] == hashlib.md5(pcms[i]).digest(), "track %d of %d, %s, %d threads" % (i, len(names), mode, threads)
            if i % 7 == 0:
                assert got == helpers.oracle_encode(pcms[i], 44100, 2, 16, o)
            os.unlink(name)
    b.lib().b200flac_pool_clear()


def test_encode_files_error_then_reuse(tmp_path, built):
    """a file that cannot be written fails the job (no hang, the other files are still produced or not -- unspecified),
    and the next job on the same cached context works"""
    b = _b200()
    o = helpers.options(block_size=4096, max_lpc_order=8, max_residual_partition_order=4)
    kw = {k: v for k, v in o.items() if k != "padding_size"}
    p = b.make_params(44100, 2, 16, **kw)
    lengths = [9000, 12000, 5000]
    pcms = [helpers.synth_pcm(900 + i, 2, 16, n) for i, n in enumerate(lengths)]
    bufs = [np.frombuffer(pcm, dtype=np.uint8).copy() for pcm in pcms]
    names = [os.path.join(str(tmp_path), "a.flac"), os.path.join(str(tmp_path), "missing_dir", "b.flac"),
             os.path.join(str(tmp_path), "c.flac")]
    with pytest.raises(Exception):
        b.encode_files(names, p, bufs, lengths, device=0, host_threads=2)
    names[1] = os.path.join(str(tmp_path), "b.flac")
    b.encode_files(names, p, bufs, lengths, device=0, host_threads=2)
    for name, pcm in zip(names, pcms):
        assert open(name, "rb").read() == helpers.oracle_encode(pcm, 44100, 2, 16, o)
    b.lib().b200flac_pool_clear()
