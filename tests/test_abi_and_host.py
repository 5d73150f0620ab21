"""CPU: the C-ABI library loads and exports every symbol include/b200flac.h declares, refuses to
work without a GPU (no CPU fallback), and the host-side sharding logic is right (incl. a
world_size-2 gloo run).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers

ROOT = helpers.ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200flac.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200flac_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    import b200flac
    lib = C.CDLL(b200flac.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libb200flac.so does not export %s" % s
    assert lib.b200flac_abi_version() == 1


def test_python_binding_covers_the_header(built):
    import b200flac
    L = b200flac.lib()
    for s in _declared_symbols():
        getattr(L, s)


def test_no_cpu_fallback(built, tmp_path):
    import b200flac
    if b200flac.device_count() > 0:
        pytest.skip("a GPU is present")
    p = b200flac.make_params()
    with pytest.raises(b200flac.B200FlacError) as e:
        b200flac.Encoder(p)
    assert "no CPU fallback" in str(e.value)
    out = os.path.join(str(tmp_path), "x.flac")
    with pytest.raises(b200flac.B200FlacError):
        b200flac.encode_file(out, p, b"\0" * 64, 16)
    assert not os.path.exists(out)


def test_struct_layouts_match_header(built):
    import b200flac
    assert C.sizeof(b200flac.Params) == 14 * 4
    assert C.sizeof(b200flac.Segment) == 24
    assert C.sizeof(b200flac.Plan) == 12 + 2 * 32


def test_product_never_references_the_oracle():
    # the product path must not import, link or execute anything under oracle/
    pkg = os.path.join(ROOT, "python-audio-tools_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "flac_oracle" not in txt and "oracle/" not in txt, \
                    os.path.join(dirpath, f)


def test_frame_ranges():
    import sharding
    for total, bs, world in ((158760000, 4096, 8), (10, 4096, 4), (4096 * 7 + 5, 4096, 2), (0, 4096, 3),
                             (691200000, 4608, 8)):
        r = sharding.frame_ranges(total, bs, world)
        assert len(r) == world
        assert sum(n for _, n, _ in r) == total
        pos = 0
        for off, n, first in r:
            if n:
                assert off == pos and off % bs == 0 and first == off // bs
            pos += n
        # only the last non-empty range may hold a short block
        nonempty = [x for x in r if x[1]]
        for off, n, _ in nonempty[:-1]:
            assert n % bs == 0
    assert [n // 4608 for _, n, _ in sharding.frame_ranges(691200000, 4608, 8)] == [18750] * 8


def _shard_worker(rank, world, port, pcm, opts, q):
    import torch.distributed as dist
    import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fb = 4
    off, n, first = sharding.frame_ranges(len(pcm) // fb, opts["block_size"], world)[rank]
    # stand-in encoder for the CPU test: the oracle encodes this rank's frame range
    frames, sizes = helpers.oracle_encode_range(pcm[off * fb:(off + n) * fb], 44100, 2, 16, opts, first)
    pcms = [min(opts["block_size"], n - i * opts["block_size"]) for i in range(len(sizes))]
    gathered = [None] * world
    dist.all_gather_object(gathered, (frames, sizes, pcms))
    if rank == 0:
        table, mn, mx, total = sharding.merge_frame_tables([g[1] for g in gathered], [g[2] for g in gathered])
        q.put((b"".join(g[0] for g in gathered), table, mn, mx, total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_frame_range_sharding_gloo(built):
    import torch.multiprocessing as mp
    opts = helpers.options(block_size=1152, max_lpc_order=8, max_residual_partition_order=4, adaptive_mid_side=True)
    pcm = helpers.synth_pcm(42, 2, 16, 1152 * 9 + 300)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, pcm, opts, q)) for r in range(2)]
    for p in procs:
        p.start()
    joined, table, mn, mx, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flac, offs = helpers.oracle_encode(pcm, 44100, 2, 16, opts, want_offsets=True)
    ff = helpers.first_frame_offset(flac)
    assert joined == flac[ff:]
    assert table == offs and total == len(flac) - ff
    si = helpers.streaminfo(flac)
    assert (mn, mx) == (si["min_frame"], si["max_frame"])


# ---------------------------------------------------------------------------------------------
# metadata finalisation in C: b200flac_finalize_metadata against the Python
# restatement of FlacAudio.from_pcm's tail (audiotools/flac.py:1811-1832 -> our audiotools/flac.py).
# Host-only: the files come from the CPU oracle.
# ---------------------------------------------------------------------------------------------
def _python_finalize(path, offsets, interval, mask):
    from audiotools import flac as aflac
    f = aflac.FlacAudio(path)
    md = f.get_metadata()
    md.add_block(f.seektable(list(offsets), interval))
    if mask:
        for i, (bid, payload) in enumerate(md.block_list):
            if bid == aflac.BLOCK_VORBIS_COMMENT:
                vlen = int.from_bytes(payload[0:4], "little")
                count = int.from_bytes(payload[4 + vlen:8 + vlen], "little")
                comment = ("WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x%.4X" % mask).encode()
                md.block_list[i] = (bid, payload[:4 + vlen] + (count + 1).to_bytes(4, "little") + payload[8 + vlen:] +
                                    len(comment).to_bytes(4, "little") + comment)
                break
    f.update_metadata(md)


@pytest.mark.parametrize("padding,mask,rate,n", [
    (4096, 0, 44100, 44100 * 25 + 17),        # three seek points, fits the padding: frames do not move
    (4096 + 4 + 3 * 18, 0x3F, 44100, 44100 * 25 + 17),   # from_pcm's own padding size, channel-mask tag
    (0, 0, 8000, 8000 * 31),                  # a zero-length PADDING cannot take the SEEKTABLE: file rewritten
    (30, 0x0603, 8000, 8000 * 45),            # padding smaller than the growth: rewritten
    (4096, 0, 44100, 0),                      # empty stream: empty SEEKTABLE
    (4096, 0, 44100, 100),                    # one short frame
])
def test_native_metadata_finalisation_matches_python(padding, mask, rate, n, tmp_path, built):
    import shutil
    import b200flac
    o = helpers.options(block_size=1152, max_lpc_order=0, max_residual_partition_order=2, padding_size=padding)
    pcm = helpers.synth_pcm(11, 1, 16, n)
    data, offsets = helpers.oracle_encode(pcm, rate, 1, 16, o, want_offsets=True)
    a, b = os.path.join(str(tmp_path), "a.flac"), os.path.join(str(tmp_path), "b.flac")
    with open(a, "wb") as fh:
        fh.write(data)
    shutil.copy(a, b)
    _python_finalize(a, offsets, rate * 10, mask)
    b200flac.finalize_metadata(b, offsets, rate * 10, mask)
    got, want = open(b, "rb").read(), open(a, "rb").read()
    assert got == want
    # frames are intact and, when the padding could absorb the SEEKTABLE, did not move
    ff = helpers.first_frame_offset(data)
    assert want.endswith(data[ff:])
    if padding >= 4096:
        assert helpers.first_frame_offset(want) == ff
    if helpers.have_ref() and n:
        assert helpers.ref_decode(got) == pcm
    # interval 0 means 10 s, like FlacAudio.seektable's default
    shutil.copy(os.path.join(str(tmp_path), "a.flac"), os.path.join(str(tmp_path), "c.flac"))


def _hand_built_flac(total_frames, padding, frames_blob, rate=100, channels=2, bps=16):
    """a FLAC file image assembled byte by byte here (no encoder, no restatement involved): fLaC, STREAMINFO,
    VORBIS_COMMENT (the encoder's vendor string, no comments), PADDING as the last block, then opaque frame bytes"""
    import struct
    si = struct.pack(">HH", 1000, 1000) + (700).to_bytes(3, "big") + (800).to_bytes(3, "big") + \
        ((rate << 44) | ((channels - 1) << 41) | ((bps - 1) << 36) | total_frames).to_bytes(8, "big") + bytes(range(16))
    vendor = b"Python Audio Tools 2.22alpha1"
    vc = struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", 0)
    return (b"fLaC" + bytes([0x00]) + (34).to_bytes(3, "big") + si + bytes([0x04]) + len(vc).to_bytes(3, "big") + vc +
            bytes([0x81]) + padding.to_bytes(3, "big") + bytes(padding) + frames_blob), si, vendor


def test_native_metadata_finalisation_hand_computed_fixture(tmp_path, built):
    """(f)1 pinned independently of this repo's Python restatement: the expected file is written out here from the
    reference's own arithmetic.  FlacAudio.seektable (audiotools/flac.py:1847-1876): offsets [(0, 1000), (700, 1000),
    (1500, 500)] give sample_offsets [0, 1000, 2000]; for pcm_frame in xrange(0, total_frames = 2500, interval = 1000)
    the points are (0, 0, 1000), (1000, 700, 1000), (2000, 1500, 500).  Flac_SEEKTABLE.build (:611-615) writes each as
    64 + 64 + 16 bits; FlacMetaData.add_block (:53-75) puts block 3 before the VORBIS_COMMENT; update_metadata
    (:1396-1424) takes the 4 + 54 bytes of growth out of the PADDING block (4096 -> 4038) and rewrites the head in
    place, so the frames do not move.  With a channel mask (from_pcm :1827-1832) the VORBIS_COMMENT gains
    "WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x003F" (4 + 40 bytes) and the padding shrinks to 3994."""
    import struct
    import b200flac
    frames = b"\xff\xf8" + bytes((i * 7 + 3) & 0xFF for i in range(2198))
    offsets = [(0, 1000), (700, 1000), (1500, 500)]
    seek = b"".join(struct.pack(">QQH", *p) for p in [(0, 0, 1000), (1000, 700, 1000), (2000, 1500, 500)])
    assert len(seek) == 54
    # --- the PADDING takes the growth ---
    data, si, vendor = _hand_built_flac(2500, 4096, frames)
    path = os.path.join(str(tmp_path), "hand.flac")
    open(path, "wb").write(data)
    b200flac.finalize_metadata(path, offsets, 1000, 0)
    vc = struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", 0)
    want = (b"fLaC" + bytes([0x00, 0, 0, 34]) + si + bytes([0x03, 0, 0, 54]) + seek +
            bytes([0x04]) + len(vc).to_bytes(3, "big") + vc + bytes([0x81]) + (4038).to_bytes(3, "big") + bytes(4038) + frames)
    assert len(want) == len(data)
    assert open(path, "rb").read() == want
    # --- the same with a channel mask: one comment more, less padding ---
    open(path, "wb").write(data)
    b200flac.finalize_metadata(path, offsets, 1000, 0x3F)
    comment = b"WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x003F"
    assert len(comment) == 40
    vc1 = struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", 1) + struct.pack("<I", 40) + comment
    want = (b"fLaC" + bytes([0x00, 0, 0, 34]) + si + bytes([0x03, 0, 0, 54]) + seek +
            bytes([0x04]) + len(vc1).to_bytes(3, "big") + vc1 + bytes([0x81]) + (3994).to_bytes(3, "big") + bytes(3994) + frames)
    assert open(path, "rb").read() == want
    # --- interval 0 = ten seconds of the STREAMINFO rate (flac.py:1856-1857): 100 Hz -> 1000 frames, the same table ---
    open(path, "wb").write(data)
    b200flac.finalize_metadata(path, offsets, 0, 0)
    assert open(path, "rb").read()[:len(want) - len(frames)][42:42 + 58] == bytes([0x03, 0, 0, 54]) + seek
    # --- a PADDING of 10 bytes cannot take 58: the file is rewritten (:1425-1462), blocks keep their sizes ---
    small, _, _ = _hand_built_flac(2500, 10, frames)
    open(path, "wb").write(small)
    b200flac.finalize_metadata(path, offsets, 1000, 0)
    want = (b"fLaC" + bytes([0x00, 0, 0, 34]) + si + bytes([0x03, 0, 0, 54]) + seek +
            bytes([0x04]) + len(vc).to_bytes(3, "big") + vc + bytes([0x81, 0, 0, 10]) + bytes(10) + frames)
    assert open(path, "rb").read() == want
    # --- a seek point between frame starts belongs to the frame that contains it (bisect_right - 1, :1869) ---
    open(path, "wb").write(data)
    b200flac.finalize_metadata(path, offsets, 900, 0)      # points at 0, 900, 1800 -> frames starting at 0, 0, 1000
    got = open(path, "rb").read()
    pts = [struct.unpack(">QQH", got[46 + 18 * i:64 + 18 * i]) for i in range(3)]
    assert pts == [(0, 0, 1000), (0, 0, 1000), (1000, 700, 1000)]


def test_native_metadata_finalisation_errors(tmp_path, built):
    import b200flac
    with pytest.raises(b200flac.B200FlacError):
        b200flac.finalize_metadata(os.path.join(str(tmp_path), "missing.flac"), [(0, 1)])
    bad = os.path.join(str(tmp_path), "bad.flac")
    with open(bad, "wb") as fh:
        fh.write(b"RIFFxxxxWAVE")
    with pytest.raises(b200flac.B200FlacError):
        b200flac.finalize_metadata(bad, [(0, 1)])


def test_standalone_driver_cli(tmp_path, built):
    """b200flacenc: the reference driver's options (flac.c:1652-1666); without a GPU it fails loudly"""
    import subprocess
    exe = os.path.join(ROOT, "python-audio-tools_b200", "b200flacenc")
    assert os.path.exists(exe)
    r = subprocess.run([exe, "--help"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0
    for flag in ("--channels", "--sample_rate", "--bits-per-sample", "--block-size", "--max-lpc-order",
                 "--min-partition-order", "--max-partition-order", "--mid-side", "--adaptive-mid-side",
                 "--exhaustive-model-search"):
        assert flag in r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "exactly 1 output file required" in r.stdout
    import b200flac
    if b200flac.device_count() == 0:
        r = subprocess.run([exe, "-q", os.path.join(str(tmp_path), "x.flac")], input=b"\0" * 64, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE)
        assert r.returncode == 1 and b"no CPU fallback" in r.stderr


def test_host_md5_lanes_match_hashlib(built):
    """csrc/md5_lanes.cpp (sixteen MD5 chains side by side, the host's share of b200flac_encode_files' hashing):
    ragged lengths around every padding boundary, fewer than sixteen strings, lanes that finish at different blocks"""
    import hashlib
    import b200flac
    rng = np.random.RandomState(11)
    for lens in ([0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 121, 127, 128, 129, 1000],
                 [4096 * 4 + 2, (1 << 20) + 3, 5], [100000 + 977 * i for i in range(16)], [64 * 1000] * 16, [7], [0]):
        datas = [rng.randint(0, 256, size=n).astype(np.uint8).tobytes() for n in lens]
        assert b200flac.host_md5_many(datas) == [hashlib.md5(d).digest() for d in datas], lens


def test_many_files_host_zone_model(built):
    """the split of b200flac_encode_files' hashing between the device and the pool (csrc/b200flac_batch.cu,
    b200flac_internal_host_zone): the pool takes the end of the list the device cannot hash in time -- pace x one device
    hash -- unless file writing leaves it no threads, and never more than it can hash by the time the device is done"""
    import b200flac
    fn = b200flac.lib().b200flac_internal_host_zone
    fn.restype = C.c_uint64
    fn.argtypes = [C.c_uint64, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double]
    track = 7938000 * 4
    hash_s = track / 100e6
    zone = lambda tracks, threads, pace=36e9: fn(tracks * track, pace, hash_s, threads, 2.0e9, 0.7, 2.5e9)
    # 1,000 and 3,000 tracks, 16 threads: what arrives during the last device hash (11.4 GB at 36 GB/s)
    assert abs(zone(1000, 16) - 36e9 * hash_s) < 1e6 and abs(zone(3000, 16) - 36e9 * hash_s) < 1e6
    # a pool of 4 threads is saturated by writing 1,000 tracks' files: no zone, every hash on the device
    assert zone(1000, 4) == 0
    # ... but has time left over on one rank's share of an 8-GPU job
    assert 0 < zone(125, 4) < 125 * track
    # a job of 16 long tracks is all zone (the device's 0.3 s per track would be the whole job)
    assert zone(16, 16) >= 16 * track
    # slower copies (eight ranks sharing the host fabric): a smaller zone; more threads: never a smaller one
    assert zone(1000, 16, pace=17e9) < zone(1000, 16)
    assert all(zone(1000, t + 1) >= zone(1000, t) for t in range(1, 32))


def test_host_md5_lanes_callback_between_pieces(built):
    """the hashing thread looks after more urgent work between pieces (b200flac_encode_files: files to write) and
    gives up when told to: the callback runs once per piece and a non-zero return ends the call with 1"""
    import hashlib
    import b200flac
    fn = b200flac.lib().b200flac_internal_md5_many
    CB = C.CFUNCTYPE(C.c_int, C.c_void_p)
    fn.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p, C.c_uint64, CB, C.c_void_p]
    datas = [bytes([i]) * (64 * 40 + i) for i in range(8)]
    ptr = (C.c_char_p * 8)(*datas)
    lens = (C.c_uint64 * 8)(*[len(d) for d in datas])
    out = (C.c_uint8 * 128)()
    calls = []
    assert fn(ptr, lens, 8, out, 64 * 10, CB(lambda arg: calls.append(1) or 0), None) == 0
    assert len(calls) == 4                                           # 40 blocks per string, 10 per piece
    assert [bytes(out[16 * i:16 * i + 16]) for i in range(8)] == [hashlib.md5(d).digest() for d in datas]
    calls = []
    assert fn(ptr, lens, 8, out, 64 * 10, CB(lambda arg: calls.append(1) or (len(calls) == 2)), None) == 1
    assert len(calls) == 2
