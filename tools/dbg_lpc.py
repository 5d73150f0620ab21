import sys, os, subprocess, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-audio-tools_b200"))
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if "torch" in mode:
    import torch
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
import b200flac
L = b200flac.lib()
proc = None
if "smi" in mode:
    proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.DEVNULL)
    time.sleep(0.5)
seed = 1235 if "seed" in mode else 1234
for nfr_pcm in (2048 * 4096, 8387820, 158760000):
    n = nfr_pcm
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)
    enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1)
    nbytes = n * 4
    cap = enc.output_bound(n, 1)
    d_pcm = L.b200flac_device_alloc(0, nbytes)
    d_out = L.b200flac_device_alloc(0, cap)
    L.b200flac_device_synth_pcm(0, d_pcm, seed, 2, 16, 0, n)
    for _ in range(4):
        enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
    k = enc.kernel_ms(0)
    print(mode, n, " ".join("%.3f" % v for v in k))
    L.b200flac_device_free(0, d_pcm); L.b200flac_device_free(0, d_out); enc.close()
if proc: proc.terminate()
