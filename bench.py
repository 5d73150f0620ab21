#!/usr/bin/env python
"""bench.py -- FLAC encode throughput of the B200 engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # this engine
    python bench.py --impl reference --steps K --warmup W  # the reference's C encoder on host cores

Workload (BASELINE.json configs[1]): 1 h of synthetic 44.1 kHz / 16-bit stereo PCM
(158,760,000 PCM frames = 317.52 M channel-samples), block_size 4096, max_lpc_order 12,
adaptive mid-side, max_residual_partition_order 6 (the reference standalone default,
src/encoders/flac.c:1647), frames sharded by range: every rank (GPU) encodes its own hour.

A step = one pass of the hot path (6 kernels: autocorrelation, Levinson/quantise, model search,
frame select, offset scan, frame pack + CRC-16) over the rank's whole hour.
  value : whole-job Msamples/s with the PCM already resident in HBM and the frames left in
          HBM (b200flac_encoder_submit_device / collect_device on two slots), wall clock around
          K steps bracketed by barrier + device synchronize, max over ranks.
  e2e   : the same metric through the call the reference's user makes -- the stream layer
          (b200flac_stream_open / write / close: what audiotools.encoders.encode_flac binds): pageable host
          PCM in, a finished .flac (STREAMINFO with MD5, frames) written to a file on tmpfs, every step;
          host<->device copies, the MD5 thread and the file write are inside the timed region.
  e2e_frame_layer : the host-buffer frame calls (b200flac_encoder_submit / collect): pinned host PCM -> H2D ->
          kernels -> D2H frame bytes, 3 batches in flight, with `copy_floor_ms` = the same bytes moved with
          no kernels (what PCIe alone costs on this box at this N).
  api   : the Python boundary (audiotools.encoders.encode_flac with a PCMReader), whole hour and many short tracks.
  configs : BASELINE.json configs[0], [2], [3], [4] resident on one GPU (N=1), byte-compared on a sample with the
          compiled reference encoder; siblings: the TTA and ALAC encoders on the same hour.
  sharded_stream : N>1 only -- ONE hour cut by frame range across the ranks, the shards concatenated on rank 0
          and compared byte for byte with the single-GPU stream.
  roofline : dominant kernel, algorithmic bytes (PCM in + frame bytes out) / its CUDA-event time; `traffic`
          from profiles/r02_traffic.json (ncu --set full of the same kernels).
  cpu_baseline : oracle/_ref/flacenc (the compiled reference) on the box's host cores, N=1 only; its output on
          the sample is also byte-compared with this engine's.

One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))

SAMPLE_RATE, CHANNELS, BPS = 44100, 2, 16
BLOCK, LPC, PO = 4096, 12, 6
HOUR_FRAMES = 158760000
METRIC = "flac_encode_msamples_per_s"
UNIT = "Msamples/s"
REF_FLACENC = os.path.join(ROOT, "oracle", "_ref", "flacenc")
ORACLE_CLI = os.path.join(ROOT, "oracle", "flacenc_oracle")


def workload_name(frames):
    return ("%.0f s synthetic 44.1 kHz/16-bit stereo per GPU, block_size 4096, max_lpc_order 12, "
            "adaptive mid-side, max_residual_partition_order 6" % (frames / SAMPLE_RATE))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(object):
    """nvidia-smi polled every 50 ms from before the warm-up to the end of the last timed region; the
    reported clocks are the samples whose timestamps fall inside a timed region (mark() brackets them),
    or, when a region is shorter than the polling period, every sample taken while the GPU was working"""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None
        self.windows = []

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            self.t_start = time.time()
        except Exception:
            self.proc = None

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_in_timed_regions": 0}
        if not self.proc:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[2]), float(f[3]), f[6:10]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if any(a - 0.03 <= r[0] <= b + 0.03 for a, b in self.windows)]
        use = inside if inside else rows
        reasons = set()
        for r in use:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if use:
            out.update(sm_mhz=statistics.median(r[1] for r in use), sm_max_mhz=max(r[2] for r in use),
                       reasons=sorted(reasons), samples=len(rows), samples_in_timed_regions=len(inside))
        return out


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own C encoder (oracle/_ref/flacenc), one process per host core
# ----------------------------------------------------------------------------------------------
def cpu_reference_round(pcm_path, nproc, sample_frames, tmpdir):
    """runs nproc reference encoders in parallel over the same sample; returns seconds"""
    flags = ["-c", str(CHANNELS), "-r", str(SAMPLE_RATE), "-b", str(BPS), "-B", str(BLOCK), "-l", str(LPC),
             "-R", str(PO), "-M"]
    t0 = time.perf_counter()
    procs = []
    for i in range(nproc):
        out = os.path.join(tmpdir, "ref%d.flac" % i)
        procs.append(subprocess.Popen([REF_FLACENC] + flags + [out], stdin=open(pcm_path, "rb"),
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference flacenc failed")
    return time.perf_counter() - t0


def make_cpu_sample(seconds, tmpdir):
    """synthetic PCM of the workload (same generator, seed 1235) written to tmpfs by the oracle CLI"""
    frames = int(seconds * SAMPLE_RATE)
    pcm_path = os.path.join(tmpdir, "sample.pcm")
    subprocess.run([ORACLE_CLI, "--synth", "1235:%d" % frames, "--dump-pcm", pcm_path, "-B", "4096", "-l", "0",
                    os.path.join(tmpdir, "discard.flac")], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    return pcm_path, frames


def cpu_baseline(seconds_per_proc=120.0, rounds=1):
    if not (os.path.exists(REF_FLACENC) and os.path.exists(ORACLE_CLI)):
        return None
    cores = os.cpu_count() or 1
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=shm) as d:
        pcm_path, frames = make_cpu_sample(seconds_per_proc, d)
        cpu_reference_round(pcm_path, min(cores, 2), frames, d)  # warm the page cache / binary
        best = min(cpu_reference_round(pcm_path, cores, frames, d) for _ in range(rounds))
        single = cpu_reference_round(pcm_path, 1, frames, d)
    samples = frames * CHANNELS
    return {"value": cores * samples / best / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d s of the workload per process, one oracle/_ref/flacenc process per core, PCM and "
                      "output on tmpfs" % seconds_per_proc,
            "single_process_value": samples / single / 1e6}


def run_reference_arm(args, rank, world, emit):
    if rank != 0:
        return
    if not (os.path.exists(REF_FLACENC) and os.path.exists(ORACLE_CLI)):
        emit({"impl": "reference", "unavailable": "oracle/_ref/flacenc not built"})
        return
    cores = os.cpu_count() or 1
    seconds = 60.0
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=shm) as d:
        pcm_path, frames = make_cpu_sample(seconds, d)
        for _ in range(args.warmup):
            cpu_reference_round(pcm_path, cores, frames, d)
        t = 0.0
        for _ in range(args.steps):
            t += cpu_reference_round(pcm_path, cores, frames, d)
    samples = frames * CHANNELS * cores * args.steps
    value = samples / t / 1e6
    sample = ("each step: %d processes (one per host core) x %d s of the workload through the compiled "
              "reference encoder" % (cores, seconds))
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32/i64/f64",
        "data": "synthetic", "config": {"workload": workload_name(HOUR_FRAMES), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


# ----------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, device resident (tier K), and their own rooflines
# ----------------------------------------------------------------------------------------------
OTHER_CONFIGS = [
    # name, sample rate, channels, bits, seconds per GPU per step, options
    ("config3 96 kHz/24-bit stereo, block 4096, lpc 12, -m -e (exhaustive), max partition order 8", 96000, 2, 24, 1800.0,
     dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True, exhaustive_model_search=True)),
    ("config4 96 kHz/24-bit 5.1 (6 ch), block 4608, lpc 12, max partition order 6", 96000, 6, 24, 600.0,
     dict(block_size=4608, max_lpc_order=12, max_residual_partition_order=6)),
    ("level8 44.1 kHz/16-bit stereo, block 4096, lpc 12, -m -e, max partition order 6 (FlacAudio.from_pcm default, config5 setting)",
     44100, 2, 16, 3600.0,
     dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True, exhaustive_model_search=True)),
]


def resident_config(b200flac, L, dev, rank, world, barrier, allmax, name, rate, ch, bps, seconds, opts, steps):
    n = int(seconds * rate)
    n -= n % opts["block_size"]
    p = b200flac.make_params(rate, ch, bps, **opts)
    enc = b200flac.Encoder(p, device=dev, max_pcm_frames_per_batch=n, n_slots=1)
    nbytes = n * ch * (bps // 8)
    cap = enc.output_bound(n, 1)
    d_pcm = L.b200flac_device_alloc(dev, nbytes)
    d_out = L.b200flac_device_alloc(dev, cap)
    if not d_pcm or not d_out:
        raise SystemExit("device allocation failed: " + L.b200flac_last_error().decode())
    L.b200flac_device_synth_pcm(dev, d_pcm, 1236 + rank, ch, bps, 0, n)
    for _ in range(3):
        enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
    launches0 = enc.launch_count()
    kms = [0.0] * 5
    dev_ms = 0.0
    barrier()
    walls = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out_bytes, nfr, ms = enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
        walls.append(time.perf_counter() - t0)
        dev_ms += ms
        for i, v in enumerate(enc.kernel_ms(0)):
            kms[i] += v
    barrier()
    # (every step is a blocking call: the median step, times the step count, so that one host hiccup -- these lines run
    # right after the API arm freed tens of GB of tmpfs and pinned memory -- does not set the rate; max over ranks)
    dt = allmax(statistics.median(walls) * steps)
    launches = enc.launch_count() - launches0
    L.b200flac_device_free(dev, d_pcm)
    L.b200flac_device_free(dev, d_out)
    enc.close()
    peak, peak_kind = peaks()
    algo = nbytes + out_bytes
    kms = [v / steps for v in kms]
    names = ["lpc_model", "analyze", "select_scan", "pack", "crc16"]
    dom = max(range(5), key=lambda i: kms[i])
    res = {"workload": name + ", %.0f s per GPU per step" % seconds,
           "value": world * n * ch * steps / dt / 1e6, "unit": UNIT, "ms_per_step": 1000.0 * dt / steps,
           "device_ms_per_step": dev_ms / steps, "best_ms_per_step": 1000.0 * min(walls),
           "device_value": n * ch / (dev_ms / steps) / 1e3,      # this rank's samples / CUDA-event time of its kernels
           "compressed_ratio": out_bytes / nbytes,
           "kernel_ms": dict(zip(names, kms)),
           "roofline": {"bound": "hbm", "kernel": names[dom], "peak": peak, "unit": "GB/s", "peak_kind": peak_kind,
                        "algorithmic_bytes_per_launch": algo,
                        "achieved": algo / (kms[dom] * 1e-3) / 1e9 if kms[dom] else None,
                        "frac": algo / (kms[dom] * 1e-3) / 1e9 / peak if kms[dom] else None,
                        "pipeline_achieved": algo / (dev_ms / steps * 1e-3) / 1e9,
                        "pipeline_frac": algo / (dev_ms / steps * 1e-3) / 1e9 / peak,
                        "bytes_per_sample": algo / (n * ch)}}
    return res, launches


def run_threads(fns):
    import threading
    errs = []

    def wrap(f):
        try:
            f()
        except BaseException as e:  # noqa: BLE001 (reported below)
            errs.append(e)
    th = [threading.Thread(target=wrap, args=(f,)) for f in fns]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise SystemExit("worker failed: %r" % (errs[0],))


def traffic_for(kernel_name, hour):
    """DRAM bytes of one launch of the dominant kernel, from the file profiles/summarize_ncu.py --json wrote
    from the round's ncu --set full capture (valid for the default 3600 s workload only)"""
    if not hour:
        return None, None
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
    except Exception:
        return None, None
    want = {"lpc_model": "k_lpc_autoc", "analyze": "k_analyze_v3", "pack": "k_pack_v3"}.get(kernel_name)
    for k, v in t.get("kernels", {}).items():
        if want and k.startswith(want):
            return v.get("dram_bytes"), "profiles/r02_traffic.json (%s, %s)" % (t.get("source"), k)
    return None, None


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="audio per GPU per step (default: the 1 h config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-api", action="store_true", help="skip the API-tier lines (single stream, encode_flac, many tracks)")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-full-check", action="store_true", help="skip decoding the whole step on the GPU after the timed region")
    ap.add_argument("--e2e-slots", type=int, default=4, help="batches in flight in the frame-layer end-to-end arm")
    ap.add_argument("--e2e-batch-blocks", type=int, default=2048, help="FLAC frames per frame-layer batch")
    ap.add_argument("--tracks", type=int, default=1000, help="3-minute tracks of the config-5 line (all ranks together)")
    ap.add_argument("--verify-seconds", type=float, default=120.0,
                    help="correctness gate inside the CPU leg (SURVEY 8d): this much of the workload is encoded to a "
                         "file through the stream layer, decoded by the compiled reference decoder and compared byte "
                         "for byte with the compiled reference encoder's file; 0 skips")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version
    # there) are sent to stderr, the result line goes to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(result_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference_arm(args, rank, world, emit)
        return

    cores = os.cpu_count() or 1
    threads = max(1, min(32, cores // world))       # host threads of this rank (its share of the box)
    os.environ.setdefault("B200FLAC_POOL", str(max(16, threads + 4)))   # idle encoders kept by the stream layer
    os.environ["B200FLAC_DEVICE"] = str(local_rank)   # the device of calls that carry no device list (encode_flac)

    import hashlib
    import numpy as np
    import torch
    import b200flac

    if not torch.cuda.is_available() or b200flac.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.set_device(local_rank)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        torch.cuda.set_device(local_rank)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        torch.cuda.set_device(local_rank)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    L = b200flac.lib()
    dev = local_rank
    n_frames_pcm = int(args.seconds * SAMPLE_RATE)
    frame_bytes = CHANNELS * BPS // 8
    pcm_bytes = n_frames_pcm * frame_bytes
    params = b200flac.make_params(SAMPLE_RATE, CHANNELS, BPS, block_size=BLOCK, max_lpc_order=LPC,
                                  max_residual_partition_order=PO, adaptive_mid_side=True)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.TemporaryDirectory(dir=shm, prefix="b200bench%d_" % rank)

    # ---- device-resident arm -----------------------------------------------------------------
    # two slots: while the kernels of step k run, the host prepares step k + 1 (frame descriptors, task lists,
    # their upload) through the asynchronous halves of the device-resident call; every step is the whole hour
    enc = b200flac.Encoder(params, device=dev, max_pcm_frames_per_batch=n_frames_pcm, n_slots=2)
    out_cap = enc.output_bound(n_frames_pcm, 1)
    d_pcm = L.b200flac_device_alloc(dev, pcm_bytes)
    d_outs = [L.b200flac_device_alloc(dev, out_cap) for _ in range(2)]
    d_out = d_outs[0]
    if not d_pcm or not all(d_outs):
        raise SystemExit("device allocation failed: " + L.b200flac_last_error().decode())
    if L.b200flac_device_synth_pcm(dev, d_pcm, 1235 + rank, CHANNELS, BPS, 0, n_frames_pcm):
        raise SystemExit("synth failed")
    segs = [(0, n_frames_pcm, 0)]

    out_bytes = 0
    clocks = ClockSampler(dev)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        out_bytes, n_flac_frames, _ = enc.encode_device(d_pcm, segs, d_out, out_cap)
        enc.encode_device(d_pcm, segs, d_outs[1], out_cap, slot=1)
    launches0 = enc.launch_count()
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for k in range(args.steps):
        enc.submit_device(d_pcm, segs, d_outs[k & 1], out_cap, slot=k & 1)
        if k:
            enc.collect_device(slot=(k - 1) & 1)
    last = enc.collect_device(slot=(args.steps - 1) & 1)
    barrier()
    elapsed = allmax(time.perf_counter() - t0)
    clocks.mark(w0, time.time())
    launches = enc.launch_count() - launches0
    d_out = d_outs[(args.steps - 1) & 1]
    out_bytes, n_flac_frames = last[0], last[1]
    samples_per_step = n_frames_pcm * CHANNELS
    value = world * samples_per_step * args.steps / elapsed / 1e6
    # kernel times: the same step alone on the device (one slot, synchronous), CUDA events recorded by the library
    # on the stream it launches on -- in the timed region above the end of one step overlaps the start of the next
    kernel_ms = [0.0] * 5
    device_ms = 0.0
    for _ in range(args.steps):
        _, _, dev_ms = enc.encode_device(d_pcm, segs, d_outs[(args.steps) & 1], out_cap)
        device_ms += dev_ms
        for i, v in enumerate(enc.kernel_ms(0)):
            kernel_ms[i] += v

    # ---- whole-step check, outside the timed region: the frames the last step left in HBM are decoded by
    # the engine's own GPU decoder (every frame header CRC-8 and frame CRC-16 checked) and the
    # PCM must equal the step's input, all of it -- the size-independent round trip at full size.  (The
    # reference decoder AND the reference encoder check a sample of the workload in the CPU leg below.) ----
    full_check = None
    if not args.no_full_check:
        info = b200flac.StreamInfo()
        info.min_block_size = info.max_block_size = BLOCK
        info.sample_rate, info.channels, info.bits_per_sample, info.total_pcm_frames = SAMPLE_RATE, CHANNELS, BPS, n_frames_pcm
        d_dec = L.b200flac_device_alloc(dev, pcm_bytes + 64)
        if not d_dec:
            raise SystemExit("device allocation failed: " + L.b200flac_last_error().decode())
        nf, kms = C.c_uint64(0), (C.c_float * 3)()
        t0 = time.perf_counter()
        if L.b200flac_decode_device(C.byref(info), d_out, out_bytes, dev, d_dec, pcm_bytes, C.byref(nf), kms):
            raise SystemExit("whole-step check failed: " + L.b200flac_last_error().decode())
        dec_s = time.perf_counter() - t0
        a = np.empty(pcm_bytes, dtype=np.uint8)
        b = np.empty(pcm_bytes, dtype=np.uint8)
        L.b200flac_device_download(dev, a.ctypes.data, d_pcm, pcm_bytes)
        L.b200flac_device_download(dev, b.ctypes.data, d_dec, pcm_bytes)
        same = bool(np.array_equal(a, b))
        del a, b
        L.b200flac_device_free(dev, d_dec)
        L.b200flac_pool_clear()   # (releases the decoder's scratch before the end-to-end arms)
        full_check = {"what": "GPU decode of the step's %d frames back to PCM, compared with the input" % nf.value,
                      "pcm_identical": same, "decode_ms": dec_s * 1e3,
                      "decode_kernel_ms": {"scan": kms[0], "frames": kms[1], "chain_emit": kms[2]}}
        if not same or nf.value != n_flac_frames:
            raise SystemExit("whole-step check failed: decoded PCM differs from the input")
    kernel_ms = [v / args.steps for v in kernel_ms]
    # CUDA-event intervals recorded by the library on the stream it launches on.  "crc16" is the separate
    # CRC kernel of the k_pack_v2 path; k_pack_v3 (the default) computes the CRC-16 inside "pack".
    names = ["lpc_model", "analyze", "select_scan", "pack", "crc16"]
    dom = max(range(5), key=lambda i: kernel_ms[i])
    traffic, traffic_source = traffic_for(names[dom], n_frames_pcm == HOUR_FRAMES)
    algo_bytes = pcm_bytes + out_bytes            # PCM in at native width + frame bytes out
    peak, peak_kind = peaks()
    achieved = algo_bytes / (kernel_ms[dom] * 1e-3) / 1e9 if kernel_ms[dom] else None
    pipeline = algo_bytes / (device_ms / args.steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_kind": peak_kind,
                "traffic_source": traffic_source,
                "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": dict(zip(names, kernel_ms)),
                "device_ms_per_step": device_ms / args.steps,
                "pipeline_achieved": pipeline, "pipeline_frac": pipeline / peak,
                "bytes_per_sample": algo_bytes / samples_per_step}

    # the hour in page-locked host memory: input of every host-buffer arm below
    h_pcm = None
    if not (args.no_e2e and args.no_api):
        h_pcm = L.b200flac_host_alloc(pcm_bytes)
        if not h_pcm:
            raise SystemExit("pinned allocation failed")
        L.b200flac_device_download(dev, h_pcm, d_pcm, pcm_bytes)
    dev1 = (C.c_int * 1)(dev)

    def encode_file(path, off_frames, n, p=params):
        if L.b200flac_encode_file(os.fsencode(path), C.byref(p), 4096, None, h_pcm + off_frames * frame_bytes, n, dev1, 1):
            raise RuntimeError(L.b200flac_last_error().decode())

    # ---- end-to-end arm (the headline): the plugin call.  b200flac_encode_file is the C-ABI entry that
    # audiotools.encoders.encode_flac wraps (stream head, frame loop, STREAMINFO MD5 on a host thread, the
    # final STREAMINFO rewrite, the file written to tmpfs).  The STREAMINFO MD5 is a serial chain per stream
    # (one core hashes ~570 MB/s), so the step's hour is cut into one stream per host thread of this rank --
    # the same cut the reference arm makes, whose encoder cannot split a stream either. ----
    e2e = None
    e2e_frame = None
    if not args.no_e2e:
        blocks = (n_frames_pcm + BLOCK - 1) // BLOCK
        per = (blocks + threads - 1) // threads * BLOCK
        parts = [(o, min(per, n_frames_pcm - o)) for o in range(0, n_frames_pcm, per)]
        paths = [os.path.join(tmp.name, "e2e_%d.flac" % i) for i in range(len(parts))]

        def e2e_stream_step():
            run_threads([(lambda i=i: encode_file(paths[i], parts[i][0], parts[i][1])) for i in range(len(parts))])

        for _ in range(2):
            e2e_stream_step()
        barrier()
        lt0 = L.b200flac_launch_count_total()
        w0 = time.time()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_stream_step()
        barrier()
        s_elapsed = allmax(time.perf_counter() - t0)
        clocks.mark(w0, time.time())
        file_bytes = sum(os.path.getsize(p) for p in paths)
        streams_frames = sum((n + BLOCK - 1) // BLOCK for _, n in parts)
        launches += L.b200flac_launch_count_total() - lt0
        # the serial floor of this arm: OpenSSL's MD5 over one stream's PCM on one core
        part0 = np.ctypeslib.as_array((C.c_uint8 * (parts[0][1] * frame_bytes)).from_address(h_pcm))
        t0 = time.perf_counter()
        hashlib.md5(part0).digest()
        md5_s = time.perf_counter() - t0
        e2e = {"value": world * samples_per_step * args.steps / s_elapsed / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": pcm_bytes, "d2h_bytes_per_step": int(file_bytes),
               "ms_per_step": 1000.0 * s_elapsed / args.steps,
               "what": "b200flac_encode_file (the C-ABI entry audiotools.encoders.encode_flac wraps: stream head, frame "
                       "loop, STREAMINFO MD5, file written to tmpfs), host PCM in; the step's hour as %d independent "
                       "streams, one per host thread of this rank (%d host cores / %d ranks)" % (len(parts), cores, world),
               "streams_per_step": len(parts), "host_threads": threads,
               "md5_floor_ms_per_step": 1000.0 * md5_s,
               "md5_floor_note": "hashlib (OpenSSL) MD5 of one stream's PCM on one core: the serial part of every stream"}

        # ---- frame-layer arm (tier D): b200flac_encoder_submit/collect = the reference's flacenc_write_frame
        # seam for batches of frames; pinned host PCM -> H2D -> kernels -> D2H frame bytes, no MD5, no file ----
        batch_frames = BLOCK * args.e2e_batch_blocks
        nslots = args.e2e_slots
        enc2 = b200flac.Encoder(params, device=dev, max_pcm_frames_per_batch=batch_frames, n_slots=nslots)
        bcap = enc2.output_bound(batch_frames, 1)
        fcap = batch_frames // BLOCK + 4
        h_out = [L.b200flac_host_alloc(bcap) for _ in range(nslots)]
        fb = [np.empty(fcap, dtype=np.uint32) for _ in range(nslots)]
        batches = []
        pos = 0
        while pos < n_frames_pcm:
            n = min(batch_frames, n_frames_pcm - pos)
            batches.append((pos, n))
            pos += n
        nb64, nf32 = C.c_uint64(0), C.c_uint32(0)

        def e2e_step():
            total = 0
            for b, (off, n) in enumerate(batches):
                slot = b % nslots
                if b >= nslots:
                    if L.b200flac_encoder_collect(enc2.h, slot, h_out[slot], bcap, C.byref(nb64),
                                                  fb[slot].ctypes.data, None, fcap, C.byref(nf32)):
                        raise SystemExit(L.b200flac_last_error().decode())
                    total += nb64.value
                seg = b200flac.Segment(0, n, off // BLOCK, 0)
                if L.b200flac_encoder_submit(enc2.h, slot, h_pcm + off * frame_bytes, C.byref(seg), 1):
                    raise SystemExit(L.b200flac_last_error().decode())
            for b in range(max(0, len(batches) - nslots), len(batches)):
                slot = b % nslots
                if L.b200flac_encoder_collect(enc2.h, slot, h_out[slot], bcap, C.byref(nb64),
                                              fb[slot].ctypes.data, None, fcap, C.byref(nf32)):
                    raise SystemExit(L.b200flac_last_error().decode())
                total += nb64.value
            return total

        for _ in range(2):
            e2e_bytes = e2e_step()
        launches_e0 = enc2.launch_count()
        barrier()
        w0 = time.time()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_bytes = e2e_step()
        barrier()
        e_elapsed = allmax(time.perf_counter() - t0)
        clocks.mark(w0, time.time())
        launches += enc2.launch_count() - launches_e0
        assert e2e_bytes == out_bytes, "frame-layer arm produced %d bytes, resident arm %d" % (e2e_bytes, out_bytes)
        for p in h_out:
            L.b200flac_host_free(p)
        enc2.close()

        # ---- the copies alone: the same bytes per step, H2D from and D2H to pinned memory, every rank at once,
        # no kernels -- the floor the host/PCIe fabric of this box sets for the frame-layer arm at this N ----
        hp = torch.empty(pcm_bytes, dtype=torch.uint8).pin_memory()
        ho = torch.empty(int(out_bytes), dtype=torch.uint8).pin_memory()
        dp = torch.empty(pcm_bytes, dtype=torch.uint8, device="cuda")
        do = torch.empty(int(out_bytes), dtype=torch.uint8, device="cuda")
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_step():
            with torch.cuda.stream(s_in):
                dp.copy_(hp, non_blocking=True)
            with torch.cuda.stream(s_out):
                ho.copy_(do, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
        copy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            copy_step()
        barrier()
        c_elapsed = allmax(time.perf_counter() - t0)
        del hp, ho, dp, do
        e2e_frame = {"value": world * samples_per_step * args.steps / e_elapsed / 1e6, "unit": UNIT,
                     "h2d_bytes_per_step": pcm_bytes, "d2h_bytes_per_step": int(e2e_bytes + 4 * n_flac_frames),
                     "ms_per_step": 1000.0 * e_elapsed / args.steps,
                     "copy_floor_ms": 1000.0 * c_elapsed / args.steps,
                     "copy_floor_what": "the same H2D and D2H bytes per rank and step as pure pinned-memory copies on two "
                                        "streams, all %d ranks at once, no kernels (max over ranks)" % world,
                     "what": "b200flac_encoder_submit/collect (frame layer, the flacenc_write_frame seam): pinned host PCM "
                             "in, frame bytes out to pinned host memory, %d-block batches, %d in flight; no MD5, no file"
                             % (batch_frames // BLOCK, nslots)}

    # ---- API tier lines (outside every timed region above) ----------------------------------------
    api = None
    if not args.no_api:
        api = {}
        # (1) one stream = the whole hour through the C entry, then through the CPython extension
        one = os.path.join(tmp.name, "one.flac")
        encode_file(one, 0, n_frames_pcm)
        barrier()
        t0 = time.perf_counter()
        encode_file(one, 0, n_frames_pcm)
        t_c = time.perf_counter() - t0
        whole = np.ctypeslib.as_array((C.c_uint8 * pcm_bytes).from_address(h_pcm))
        t0 = time.perf_counter()
        hashlib.md5(whole).digest()
        t_md5 = time.perf_counter() - t0
        one_sha = hashlib.sha256(open(one, "rb").read()).hexdigest()
        api["single_stream_encode_file"] = {
            "value": samples_per_step / t_c / 1e6, "unit": UNIT, "ms": 1000.0 * t_c,
            "md5_alone_ms": 1000.0 * t_md5, "md5_alone_value": samples_per_step / t_md5 / 1e6,
            "what": "b200flac_encode_file, the step's whole hour as ONE stream on one GPU (host PCM -> file on tmpfs); "
                    "md5_alone = hashlib MD5 of the same PCM on one core, the floor of any single stream"}
        try:
            import audiotools
            from audiotools import encoders as at_enc
            rd = audiotools.PCMBytesReader(whole, SAMPLE_RATE, CHANNELS, 0x3, BPS)
            py = os.path.join(tmp.name, "py.flac")
            t0 = time.perf_counter()
            offs = at_enc.encode_flac(py, audiotools.BufferedPCMReader(rd), BLOCK, LPC, 0, PO, adaptive_mid_side=1)
            t_py = time.perf_counter() - t0
            same = hashlib.sha256(open(py, "rb").read()).hexdigest() == one_sha
            api["single_stream_encode_flac"] = {
                "value": samples_per_step / t_py / 1e6, "unit": UNIT, "ms": 1000.0 * t_py, "frames": len(offs),
                "identical_to_encode_file": bool(same),
                "what": "audiotools.encoders.encode_flac(filename, BufferedPCMReader(reader), 4096, 12, 0, 6, "
                        "adaptive_mid_side=1) -- the entry point north_star names -- over a Python PCMReader handing out "
                        "FrameLists, same hour, same GPU"}
            if not same:
                raise SystemExit("encode_flac and b200flac_encode_file wrote different files")
        except ImportError as e:
            api["single_stream_encode_flac"] = {"unavailable": "audiotools extension not importable: %s" % e}
        del whole
        # (2) config 5's shape: many 3-minute tracks at FlacAudio.from_pcm's default level 8, one host thread per
        # core of this rank, every rank on its own GPU
        tn = 7938000
        if n_frames_pcm >= tn:
            p8 = b200flac.make_params(SAMPLE_RATE, CHANNELS, BPS, block_size=4096, max_lpc_order=12,
                                      max_residual_partition_order=6, mid_side=True, exhaustive_model_search=True)
            distinct = n_frames_pcm // tn             # distinct track PCMs cut from the hour
            my_tracks = max(threads, args.tracks // world)
            nxt = [0]
            import threading
            lock = threading.Lock()

            def track_worker(tid):
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= my_tracks:
                        return
                    encode_file(os.path.join(tmp.name, "trk_%d.flac" % tid), (i % distinct) * tn, tn, p8)

            nxt[0] = max(0, my_tracks - 2 * threads)  # warm-up: fills the encoder pool for these options
            run_threads([(lambda t=t: track_worker(t)) for t in range(threads)])
            nxt[0] = 0
            barrier()
            t0 = time.perf_counter()
            run_threads([(lambda t=t: track_worker(t)) for t in range(threads)])
            barrier()
            t_tr = allmax(time.perf_counter() - t0)
            total_tracks = allsum(my_tracks)
            api["many_tracks"] = {
                "value": total_tracks * tn * CHANNELS / t_tr / 1e6, "unit": UNIT, "tracks": int(total_tracks),
                "tracks_per_s": total_tracks / t_tr, "seconds": t_tr, "host_threads_per_rank": threads,
                "projected_s_for_10000_tracks": 10000.0 * t_tr / total_tracks,
                "what": "BASELINE config 5's shape: %d three-minute 44.1 kHz/16-bit stereo tracks (%d distinct PCMs cut from "
                        "the hour, cycled) through b200flac_encode_file at level 8 (-m -e, lpc 12, the from_pcm default), "
                        "%d host threads per rank, %d GPU(s), files on tmpfs" % (int(total_tracks), distinct, threads, world)}
            # (3) the same tracks as ONE job per rank: b200flac_encode_files packs them into many-segment batches and
            # hashes every track's MD5 on the device; files are byte-identical to (2)'s (checked on a sample)
            names = [os.path.join(tmp.name, "bt_%d.flac" % i) for i in range(my_tracks)]
            c_names = (C.c_char_p * my_tracks)(*[os.fsencode(x) for x in names])
            c_ptrs = (C.c_void_p * my_tracks)(*[h_pcm + (i % distinct) * tn * frame_bytes for i in range(my_tracks)])
            c_lens = (C.c_uint64 * my_tracks)(*([tn] * my_tracks))

            def batch_job(k):
                if L.b200flac_encode_files(k, c_names, C.byref(p8), 4096, None, c_ptrs, c_lens, dev, threads):
                    raise SystemExit("b200flac_encode_files failed: " + L.b200flac_last_error().decode())
            batch_job(my_tracks)                    # warm-up: the device ring (one region per batch in flight), pinned buffers, encoder
            barrier()
            t0 = time.perf_counter()
            batch_job(my_tracks)
            barrier()
            t_bt = allmax(time.perf_counter() - t0)
            st = (C.c_double * 6)()
            L.b200flac_internal_batch_stats(st)
            encode_file(os.path.join(tmp.name, "bt_check.flac"), ((my_tracks - 1) % distinct) * tn, tn, p8)
            same = open(os.path.join(tmp.name, "bt_check.flac"), "rb").read() == open(names[my_tracks - 1], "rb").read()
            if not same:
                raise SystemExit("b200flac_encode_files and b200flac_encode_file wrote different files")
            for x in names:
                os.unlink(x)
            api["many_tracks_batch"] = {
                "value": total_tracks * tn * CHANNELS / t_bt / 1e6, "unit": UNIT, "tracks": int(total_tracks),
                "tracks_per_s": total_tracks / t_bt, "seconds": t_bt, "host_threads_per_rank": threads,
                "projected_s_for_10000_tracks": 10000.0 * t_bt / total_tracks, "identical_to_encode_file": True,
                "speedup_over_one_call_per_file": t_tr / t_bt,
                "hashed_rank0": {"device_tracks": int(st[0]), "host_tracks": int(st[1]),
                                 "host_hash_mb_per_s_per_thread": st[2] / 1e6, "file_write_mb_per_s_per_thread": st[3] / 1e6,
                                 "host_hash_thread_s": st[4], "file_write_thread_s": st[5]},
                "what": "the same tracks as ONE b200flac_encode_files job per rank: many-segment batches of the frame "
                        "layer, the tracks' STREAMINFO MD5s computed on the device (one thread per track, whole batches from the "
                        "front of the list) and by spare pool threads (the end of the list, sixteen tracks at a time in vector lanes), %d host "
                        "threads write the files to tmpfs; PCM read from page-locked memory" % threads}

    # ---- the other BASELINE configurations, device resident ----
    configs = None
    if not args.no_configs:
        if h_pcm:
            L.b200flac_host_free(h_pcm)
            h_pcm = None
        L.b200flac_pool_clear()
        configs = []
        for (name, rate, ch, bps, seconds, opts) in OTHER_CONFIGS:
            res, ln = resident_config(b200flac, L, dev, rank, world, barrier, allmax, name, rate, ch, bps, seconds, opts,
                                      max(3, min(args.steps, 7)))
            launches += ln
            configs.append(res)

    # ---- the sibling encoders of SURVEY 8(f)-4 on the same runtime: TTA and ALAC, device resident, rank 0 ----
    siblings = None
    if rank == 0 and not args.no_configs:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import alac_perf
            import tta_perf
            peak, peak_kind = peaks()
            siblings = {}
            for name, mod, secs in (("tta", tta_perf, 3600.0), ("alac", alac_perf, 3600.0)):
                r = mod.measure(secs, device=dev)
                r["what"] = ("%s encode of %.0f s synthetic 44.1 kHz/16-bit stereo, PCM resident in HBM, frames left in HBM; "
                             "value = samples / CUDA-event time of its kernels" % (name.upper(), secs))
                r["roofline"] = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_kind": peak_kind,
                                 "achieved": r["algorithmic_bytes"] / (r["kernels_ms"] * 1e-3) / 1e9,
                                 "frac": r["algorithmic_bytes"] / (r["kernels_ms"] * 1e-3) / 1e9 / peak,
                                 "note": "latency-bound by construction: both codecs adapt their predictor and their Rice/Golomb "
                                         "state after every sample, so a channel of a frame is one serial chain per thread"}
                if world == 1 and not args.no_cpu_baseline:
                    r["cpu_baseline"] = mod.reference_single_core()
                siblings[name] = r
            launches += 3 * 3 + 7 * 3
        except Exception as e:          # (reported, not fatal: the FLAC line is the bench)
            siblings = {"error": repr(e)}

    # ---- N > 1: ONE stream, frame-range sharded over the N GPUs by the stream layer (config 4's shape) ----
    sharded = None
    if world > 1 and not args.no_api:
        barrier()
        done_flag = os.path.join(tempfile.gettempdir(), "b200bench_sharded_%s" % os.environ.get("MASTER_PORT", "0"))
        if rank == 0:
            try:
                rate, ch, bps, secs = 96000, 6, 24, 120.0
                o4 = dict(block_size=4608, max_lpc_order=12, max_residual_partition_order=6)
                n4 = int(secs * rate) // 4608 * 4608 + 1000
                p4 = b200flac.make_params(rate, ch, bps, **o4)
                pcm4 = b200flac.synth_pcm(1238, ch, bps, n4, device=dev)
                res = {}
                for label, devs in (("one", [dev]), ("all", list(range(world)))):
                    path = os.path.join(tmp.name, "shard_%s.flac" % label)
                    b200flac.encode_file(path, p4, pcm4, n4, devices=devs)        # warm (encoder pool)
                    t0 = time.perf_counter()
                    b200flac.encode_file(path, p4, pcm4, n4, devices=devs)
                    res[label] = (time.perf_counter() - t0, hashlib.sha256(open(path, "rb").read()).hexdigest())
                sharded = {"what": "one %d s 96 kHz/24-bit 6-channel stream (config 4's shape, block 4608) through "
                                   "b200flac_encode_file with devices = all %d GPUs: the stream layer hands batches of "
                                   "consecutive frames to the devices in turn and concatenates their frames in order on "
                                   "the host; compared with the same call on one GPU" % (secs, world),
                           "n_devices": world, "identical_to_one_gpu": res["one"][1] == res["all"][1],
                           "sha256": res["all"][1], "ms_one_gpu": 1000.0 * res["one"][0], "ms_all_gpus": 1000.0 * res["all"][0],
                           "value_all_gpus": n4 * ch / res["all"][0] / 1e6, "unit": UNIT,
                           "note": "a single stream is bound by its serial STREAMINFO MD5 on one host core, not by the GPUs"}
            finally:
                open(done_flag, "w").close()
            if sharded and not sharded["identical_to_one_gpu"]:
                raise SystemExit("sharded stream differs from the single-GPU stream")
        else:
            while not os.path.exists(done_flag):    # (no collective here: a waiting NCCL kernel would sit on the GPUs)
                time.sleep(0.05)
        barrier()
        if rank == 0:
            try:
                os.unlink(done_flag)
            except OSError:
                pass

    clk = clocks.stop()

    # ---- CPU leg (rank 0, N = 1): the compiled reference as baseline and as checker.  The only place this
    # arm executes anything under oracle/: the reference encoder is timed, and for a sample of the workload
    # encoded through the stream layer (a) the reference decoder (CRC-16 of every frame, STREAMINFO MD5) must
    # return the input PCM and (b) the file must equal the reference encoder's, byte for byte -- SURVEY 8(d)'s
    # correctness gate, outside the timed regions ----
    base = None
    verified = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline()
        REF_FLACDEC = os.path.join(ROOT, "oracle", "_ref", "flacdec")
        if args.verify_seconds > 0 and os.path.exists(REF_FLACDEC):
            vn = min(n_frames_pcm, int(args.verify_seconds * SAMPLE_RATE))
            host = np.empty(vn * frame_bytes, dtype=np.uint8)
            L.b200flac_device_download(dev, host.ctypes.data, d_pcm, vn * frame_bytes)
            with tempfile.TemporaryDirectory(dir=shm) as vd:
                vpath = os.path.join(vd, "v.flac")
                b200flac.encode_file(vpath, params, host, vn)
                r = subprocess.run([REF_FLACDEC, vpath], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                ok = r.returncode == 0 and r.stdout == host.tobytes()
                ppath, rpath = os.path.join(vd, "v.pcm"), os.path.join(vd, "ref.flac")
                host.tofile(ppath)
                flags = ["-c", str(CHANNELS), "-r", str(SAMPLE_RATE), "-b", str(BPS), "-B", str(BLOCK), "-l", str(LPC),
                         "-R", str(PO), "-M"]
                rr = subprocess.run([REF_FLACENC] + flags + [rpath], stdin=open(ppath, "rb"),
                                    stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                ours, theirs = open(vpath, "rb").read(), open(rpath, "rb").read() if rr.returncode == 0 else b""
                verified = {"decoder": "oracle/_ref/flacdec (reference src/decoders/flac.c)", "seconds": vn / SAMPLE_RATE,
                            "lossless": bool(ok), "file_bytes": len(ours),
                            "encoder": "oracle/_ref/flacenc (reference src/encoders/flac.c) on the same PCM and options",
                            "reference_file_bytes": len(theirs), "identical_to_reference_file": bool(ours == theirs),
                            "size_ratio_to_reference": len(ours) / len(theirs) if theirs else None}
            if not ok:
                raise SystemExit("correctness gate failed: the reference decoder did not return the input PCM")
            if ours != theirs:
                raise SystemExit("correctness gate failed: the file differs from the reference encoder's")
        if base is not None:
            base["checked"] = verified

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1000.0 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32/i64/f64",
            "data": "synthetic",
            "config": {"workload": workload_name(n_frames_pcm), "block_size": BLOCK, "max_lpc_order": LPC,
                       "max_residual_partition_order": PO, "adaptive_mid_side": True,
                       "flac_frames_per_step_per_gpu": int(n_flac_frames),
                       "compressed_ratio": out_bytes / pcm_bytes,
                       "l2": "inputs (%.0f MB per step) larger than the 126 MB L2" % (pcm_bytes / 1e6),
                       "sharding": "frame range per GPU, no collective"},
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "e2e_frame_layer": e2e_frame,
            "api": api, "configs": configs, "siblings": siblings, "sharded_stream": sharded, "gpu_launches": int(launches),
            "full_check": full_check,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"],
                       "samples": clk["samples"], "samples_in_timed_regions": clk["samples_in_timed_regions"]},
        }
        emit(line)
    if h_pcm:
        L.b200flac_host_free(h_pcm)
    L.b200flac_device_free(dev, d_pcm)
    for p_ in d_outs:
        L.b200flac_device_free(dev, p_)
    enc.close()
    tmp.cleanup()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
