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
