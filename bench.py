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
          HBM (b200flac_encoder_encode_device), wall clock around K steps bracketed by
          barrier + device synchronize, max over ranks.
  e2e   : same metric through the host-buffer C-ABI calls (b200flac_encoder_submit/collect):
          pinned host PCM -> H2D -> kernels -> D2H frame bytes, every step, 3 batches in flight.
  roofline : dominant kernel, algorithmic bytes (PCM in + frame bytes out) / its CUDA-event time.
  cpu_baseline : oracle/_ref/flacenc (the compiled reference) on the box's host cores, N=1 only.

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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="audio per GPU per step (default: the 1 h config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-check", action="store_true", help="skip decoding the whole step on the GPU after the timed region")
    ap.add_argument("--e2e-slots", type=int, default=4, help="batches in flight in the end-to-end arm")
    ap.add_argument("--e2e-batch-blocks", type=int, default=2048, help="FLAC frames per end-to-end batch")
    ap.add_argument("--verify-seconds", type=float, default=120.0,
                    help="correctness gate inside the CPU leg (SURVEY 8d): this much of the workload is encoded to a "
                         "file through the stream layer and decoded by the compiled reference decoder; 0 skips")
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
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    L = b200flac.lib()
    dev = local_rank
    n_frames_pcm = int(args.seconds * SAMPLE_RATE)
    frame_bytes = CHANNELS * BPS // 8
    pcm_bytes = n_frames_pcm * frame_bytes
    params = b200flac.make_params(SAMPLE_RATE, CHANNELS, BPS, block_size=BLOCK, max_lpc_order=LPC,
                                  max_residual_partition_order=PO, adaptive_mid_side=True)

    # ---- device-resident arm -----------------------------------------------------------------
    enc = b200flac.Encoder(params, device=dev, max_pcm_frames_per_batch=n_frames_pcm, n_slots=1)
    out_cap = enc.output_bound(n_frames_pcm, 1)
    d_pcm = L.b200flac_device_alloc(dev, pcm_bytes)
    d_out = L.b200flac_device_alloc(dev, out_cap)
    if not d_pcm or not d_out:
        raise SystemExit("device allocation failed: " + L.b200flac_last_error().decode())
    if L.b200flac_device_synth_pcm(dev, d_pcm, 1235 + rank, CHANNELS, BPS, 0, n_frames_pcm):
        raise SystemExit("synth failed")
    segs = [(0, n_frames_pcm, 0)]

    out_bytes = 0
    clocks = ClockSampler(dev)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        out_bytes, n_flac_frames, _ = enc.encode_device(d_pcm, segs, d_out, out_cap)
    launches0 = enc.launch_count()
    barrier()
    w0 = time.time()
    kernel_ms = [0.0] * 5
    device_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_bytes, n_flac_frames, dev_ms = enc.encode_device(d_pcm, segs, d_out, out_cap)
        device_ms += dev_ms
        for i, v in enumerate(enc.kernel_ms(0)):
            kernel_ms[i] += v
    barrier()
    elapsed = time.perf_counter() - t0
    clocks.mark(w0, time.time())
    launches = enc.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    samples_per_step = n_frames_pcm * CHANNELS
    value = world * samples_per_step * args.steps / elapsed / 1e6

    # ---- whole-step check, outside the timed region: the frames the last step left in HBM are decoded by
    # the engine's own GPU decoder (every frame header CRC-8 and frame CRC-16 checked) and the
    # PCM must equal the step's input, all of it -- the size-independent round trip at full size.  (The
    # reference decoder checks a sample of the workload in the CPU leg below.) ----
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
        L.b200flac_pool_clear()   # (releases the decoder's scratch before the end-to-end arm)
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
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload, from the ncu --set full
    # capture summarised in profiles/r01_v7_summary.txt (only valid for the default 3600 s workload)
    traffic_ncu = {"lpc_model": 636.3e6 + 14.6e6, "analyze": 673.2e6 + 20.6e6, "pack": 655.5e6 + 403.7e6}
    traffic = traffic_ncu.get(names[dom]) if n_frames_pcm == HOUR_FRAMES else None
    algo_bytes = pcm_bytes + out_bytes            # PCM in at native width + frame bytes out
    peak, peak_kind = peaks()
    achieved = algo_bytes / (kernel_ms[dom] * 1e-3) / 1e9
    pipeline = algo_bytes / (sum(kernel_ms) * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                "traffic_source": "profiles/r01_v7_summary.txt (ncu --set full, bytes per launch)" if traffic else None,
                "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": dict(zip(names, kernel_ms)),
                "device_ms_per_step": device_ms / args.steps,
                "pipeline_achieved": pipeline, "pipeline_frac": pipeline / peak,
                "bytes_per_sample": algo_bytes / samples_per_step}

    # ---- end-to-end arm: pinned host PCM -> H2D -> kernels -> D2H, through submit/collect ---------
    e2e = None
    if not args.no_e2e:
        batch_frames = BLOCK * args.e2e_batch_blocks
        nslots = args.e2e_slots
        enc2 = b200flac.Encoder(params, device=dev, max_pcm_frames_per_batch=batch_frames, n_slots=nslots)
        h_pcm = L.b200flac_host_alloc(pcm_bytes)
        if not h_pcm:
            raise SystemExit("pinned allocation failed")
        L.b200flac_device_download(dev, h_pcm, d_pcm, pcm_bytes)
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
        e_elapsed = time.perf_counter() - t0
        clocks.mark(w0, time.time())
        launches += enc2.launch_count() - launches_e0
        if dist is not None:
            t = torch.tensor([e_elapsed], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_elapsed = float(t.item())
        assert e2e_bytes == out_bytes, "e2e arm produced %d bytes, resident arm %d" % (e2e_bytes, out_bytes)
        e2e = {"value": world * samples_per_step * args.steps / e_elapsed / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": pcm_bytes, "d2h_bytes_per_step": int(e2e_bytes + 4 * n_flac_frames),
               "ms_per_step": 1000.0 * e_elapsed / args.steps,
               "what": "b200flac_encoder_submit/collect, pinned host PCM in, frame bytes out to pinned host "
                       "memory, %d-block batches, %d in flight" % (batch_frames // BLOCK, nslots)}
        for p in h_out:
            L.b200flac_host_free(p)
        L.b200flac_host_free(h_pcm)
        enc2.close()

    clk = clocks.stop()

    # ---- CPU leg (rank 0, N = 1): the compiled reference as baseline and as checker.  The only place this
    # arm executes anything under oracle/: the reference encoder is timed, and the reference decoder
    # (CRC-16 of every frame, STREAMINFO MD5) must return the input PCM for a sample of the workload
    # encoded through the stream layer -- SURVEY 8(d)'s correctness gate, outside the timed regions ----
    base = None
    verified = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline()
        REF_FLACDEC = os.path.join(ROOT, "oracle", "_ref", "flacdec")
        if args.verify_seconds > 0 and os.path.exists(REF_FLACDEC):
            vn = min(n_frames_pcm, int(args.verify_seconds * SAMPLE_RATE))
            host = np.empty(vn * frame_bytes, dtype=np.uint8)
            L.b200flac_device_download(dev, host.ctypes.data, d_pcm, vn * frame_bytes)
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
            with tempfile.TemporaryDirectory(dir=shm) as vd:
                vpath = os.path.join(vd, "v.flac")
                b200flac.encode_file(vpath, params, host, vn)
                r = subprocess.run([REF_FLACDEC, vpath], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                ok = r.returncode == 0 and r.stdout == host.tobytes()
                verified = {"decoder": "oracle/_ref/flacdec (reference src/decoders/flac.c)", "seconds": vn / SAMPLE_RATE,
                            "lossless": bool(ok), "file_bytes": os.path.getsize(vpath)}
            if not ok:
                raise SystemExit("correctness gate failed: the reference decoder did not return the input PCM")
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
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "gpu_launches": int(launches),
            "full_check": full_check,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"],
                       "samples": clk["samples"], "samples_in_timed_regions": clk["samples_in_timed_regions"]},
        }
        emit(line)
    L.b200flac_device_free(dev, d_pcm)
    L.b200flac_device_free(dev, d_out)
    enc.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
