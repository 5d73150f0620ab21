// b200flac_encoder.cu -- frame layer of the C ABI (include/b200flac.h): one
// encoder per CUDA device, batches of independent frames through six kernels
// (k_lpc_autoc, k_lpc_finish, k_analyze_v3 [+ k_analyze_v2 for the shapes v3 does not take],
// k_frame_select, k_scan_offsets, k_pack_v3 [or k_zero_output + k_pack_v2 + k_frame_crc16]).
//
// Replaces, for a whole batch at a time, the reference's per-frame call
//   flacenc_write_frame(encoder.frame, &encoder, samples)   src/encoders/flac.c:258, 520-671
// No CPU fallback exists: without a usable CUDA device every call fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <vector>

#include "flac_types.h"
#include "flac_common.cuh"
#include "k_lpc_model.cuh"
#include "k_analyze.cuh"
#include "k_pack.cuh"
#include "k_analyze_v2.cuh"
#include "k_analyze_v3.cuh"
#include "k_pack_v2.cuh"
#include "k_pack_v3.cuh"
#include "k_synth.cuh"

// opt-in dynamic shared memory ceiling set on every kernel that may need more than 48 KB.  The attribute
// belongs to the kernel, not to an encoder: setting each encoder's own size let one encoder lower the
// limit under another (concurrent or pooled) encoder with a larger block size.
#define BF_SMEM_OPTIN (200 * 1024)
#define BF_MAX_SEGMENTS 65536
#define BF_NUM_EVENTS 6
#define BF_MAX_CHUNKS 64
#define BF_MAX_MODEL_STREAMS 8

// A batch runs as a pipeline of chunks: contiguous frame ranges, each a mini-batch of its own (every per-frame
// and per-unit array is addressed from the chunk's first frame; task and odd-frame lists are chunk-relative).
struct Chunk {
    u32 frame0, n_frames;
    u32 task0, n_tasks;
    u32 odd0, n_odd;
};

static thread_local char g_err[512] = "";

static void set_err(const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
}

extern "C" void b200flac_internal_set_error(const char* msg) { set_err(msg); }

#define CU_CHECK(call, ret)                                                              \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf(g_err, sizeof(g_err), "%s failed: %s (%s:%d)", #call,               \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                        \
            return ret;                                                                  \
        }                                                                                \
    } while (0)

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[BF_NUM_EVENTS] = {};
    uint8_t* h_pcm = nullptr;      // pinned
    uint8_t* h_out = nullptr;      // pinned (b200flac_encoder_slot_out), out_cap bytes
    uint8_t* d_pcm = nullptr;
    bf_frame_desc* h_fd = nullptr; // pinned
    bf_frame_desc* d_fd = nullptr;
    double* h_win = nullptr;       // pinned
    double* d_win = nullptr;
    size_t win_cap = 0;            // doubles
    bf_lpc_head* d_heads = nullptr;
    double* d_autoc = nullptr;     // [units][maxl + 1] lag sums (k_lpc_autoc -> k_lpc_finish)
    uint8_t* d_lpc_wasted = nullptr;
    u32* d_ticket = nullptr;       // task counter of the persistent autocorrelation kernel
    bf_lpc_task* h_tasks = nullptr; // pinned: runs of equal-length frames (one warp task each)
    bf_lpc_task* d_tasks = nullptr;
    u32 n_tasks = 0;
    u32* h_odd = nullptr;          // pinned: frames whose length is not block_size (k_analyze_v2 takes them)
    u32* d_odd = nullptr;
    u32 n_odd = 0;
    short* d_coefs = nullptr;
    b200flac_plan* d_plans = nullptr;
    uint8_t* d_rice = nullptr;
    bf_frame_choice* d_choice = nullptr;
    u32* d_frame_bytes = nullptr;
    u64* d_frame_off = nullptr;
    u64* d_total = nullptr;
    uint8_t* d_out = nullptr;
    u32* h_frame_bytes = nullptr;  // pinned
    u64* h_total = nullptr;        // pinned
    int* d_gsamples = nullptr;
    u64* d_gheap = nullptr;
    uint8_t* d_gkarr = nullptr;
    u32 n_frames = 0;
    u64 dev_cap = 0;               // capacity of the caller's output buffer (device-resident batches)
    bool busy = false;
    bool timed = false;
    std::vector<u32> frame_pcm;
    // the software pipeline over chunks of a batch (see launch_batch)
    cudaStream_t stream_b = nullptr;     // second analysis/packing stream (odd chunks)
    cudaStream_t stream_hi[BF_MAX_MODEL_STREAMS] = {};   // streams of the floating-point model kernels
    cudaStream_t stream_lo[BF_MAX_MODEL_STREAMS] = {};   // ... the same at normal priority (tuning)
    cudaEvent_t ev_done = nullptr;       // blocking-sync event: a host thread waiting for the slot sleeps instead of spinning
    cudaEvent_t ev_in = nullptr, ev_join = nullptr;
    cudaEvent_t ev_model[BF_MAX_CHUNKS] = {}, ev_an[BF_MAX_CHUNKS] = {}, ev_scan[BF_MAX_CHUNKS] = {};
    u64* d_chunk_tot = nullptr;          // running output size after each chunk
    std::vector<Chunk> chunks;
};

struct b200flac_encoder {
    b200flac_params params;
    bf_dev_params P;
    int device;
    u64 max_pcm_frames;
    u32 max_frames;       // frames per batch the buffers are sized for
    u64 out_cap;          // bytes of d_out per slot
    int n_slots;
    Slot* slots;
    int S;                // samples per thread (template selector)
    int NT;               // threads per CTA for analyze/pack
    size_t smem_analyze, smem_pack;
    bool fast;            // the shared-memory resident kernels (k_analyze_v2 / k_pack_v2) apply
    u32 stage_words;      // shared-memory image of one subframe, in words (k_pack_v2)
    bool v3;              // k_analyze_v3 applies to the full-length blocks
    bool p3;              // k_pack_v3 (whole frame per CTA, CRC-16 fused) applies
    u32 p3_img_words;
    size_t p3_smem;
    unsigned short* d_crc_tab;   // [256] byte table + [69 + n] powers of x (see k_pack_v3)
    u32 v3_S, v3_F, v3_NT, v3_sub;   // v3_sub: run sums per thread run (2: partition order 8)
    size_t v3_smem;
    int v3_occ;           // resident CTAs per SM of k_analyze_v3 (the kernel is persistent)
    std::map<u32, std::vector<double>>* windows;
    u64 launches;
    int lpc_occ[2];       // resident one-warp CTAs per SM of k_lpc_autoc (G = 1, 2)
    int n_sms;
    u32 chunk_frames;     // frames per pipeline chunk (0: a batch is one chunk, kernels back to back on one stream)
    u32 lookahead;        // chunks the model streams may run ahead of the analysis
    u32 model_streams;    // streams the model kernels of consecutive chunks rotate over
    u32 model_priority;   // 1: those streams have high priority
    u32 lpc_grid_cap;     // pipeline: at most this many one-warp CTAs per model launch (0: one per task)
};

// kernel launches of this process (every encoder, every thread): what bench.py reports as gpu_launches
static std::atomic<unsigned long long> g_launches_total(0);
static inline void count_launches(b200flac_encoder* enc, u64 k)
{
    enc->launches += k;
    g_launches_total.fetch_add(k, std::memory_order_relaxed);
}
extern "C" uint64_t b200flac_launch_count_total(void) { return g_launches_total.load(); }

// Wait for a slot's event: poll for a short while (a batch in a running pipeline is usually done or nearly so, and
// waking a sleeping thread costs more than the wait), then sleep on the blocking-sync event, so that the many
// host threads of the many-streams case do not spin against the streams' MD5 threads.
static cudaError_t wait_event(cudaEvent_t ev)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) return e;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(1000)) break;
    }
    return cudaEventSynchronize(ev);
}

// ---- derived options ------------------------------------------------------
static u32 sample_rate_code(u32 sr) // flac.c:452-476
{
    switch (sr) {
    case 88200: return 1; case 176400: return 2; case 192000: return 3; case 8000: return 4;
    case 16000: return 5; case 22050: return 6; case 24000: return 7; case 32000: return 8;
    case 44100: return 9; case 48000: return 10; case 96000: return 11;
    default:
        if (sr <= 255000 && sr % 1000 == 0) return 0xC;
        if (sr <= 655350 && sr % 10 == 0) return 0xE;
        if (sr <= 0xFFFF) return 0xD;
        return 0;
    }
}

static u32 bps_code(u32 bps) // flac.c:479-486
{
    switch (bps) { case 8: return 1; case 12: return 2; case 16: return 4; case 20: return 5; case 24: return 6; default: return 0; }
}

// Tukey window exactly as flac.c:1139-1161 computes it (host libm cos, same as the reference)
static void tukey_window(u32 N, double* w)
{
    const double alpha = 0.5;
    const unsigned window1 = (unsigned)(alpha * (N - 1)) / 2;
    const unsigned window2 = (unsigned)((N - 1) * (1.0 - (alpha / 2.0)));
    for (unsigned n = 0; n < N; n++) {
        if (n <= window1) w[n] = 0.5 * (1.0 + cos(M_PI * (((2 * n) / (alpha * (N - 1))) - 1.0)));
        else if (n <= window2) w[n] = 1.0;
        else w[n] = 0.5 * (1.0 + cos(M_PI * (((2.0 * n) / (alpha * (N - 1))) - (2.0 / alpha) + 1.0)));
    }
}

static void derive_params(const b200flac_params* p, bf_dev_params* P)
{
    memset(P, 0, sizeof(*P));
    P->block_size = p->block_size;
    P->max_lpc_order = p->max_lpc_order;
    P->po_lim = p->max_residual_partition_order > BF_MAX_PO ? BF_MAX_PO : p->max_residual_partition_order;
    P->channels = p->channels;
    P->bps = p->bits_per_sample;
    P->bytes_ps = p->bits_per_sample / 8;
    P->sample_rate = p->sample_rate;
    const u32 bs = p->block_size; // flac.c:165-178
    P->precision = bs <= 192 ? 7 : bs <= 384 ? 8 : bs <= 576 ? 9 : bs <= 1152 ? 10 : bs <= 2304 ? 11 : bs <= 4608 ? 12 : 13;
    P->max_rice = p->bits_per_sample <= 16 ? 0xE : 0x1E; // flac.c:180-184
    P->stereo = (p->channels == 2 && (p->mid_side || p->adaptive_mid_side)) ? 1 : 0; // flac.c:532-533
    P->K = P->stereo ? 4 : p->channels;
    P->mid_side = p->mid_side ? 1 : 0;
    P->exhaustive = p->exhaustive_model_search ? 1 : 0;
    P->try_verbatim = !p->no_verbatim_subframes; // flac.c:681-685
    P->try_constant = !p->no_constant_subframes;
    P->try_fixed = !p->no_fixed_subframes;
    P->try_lpc = !(p->no_lpc_subframes || p->max_lpc_order == 0);
    P->rice_stride = 1u << P->po_lim;
    const u32 L = p->max_lpc_order;
    P->model_stride = L ? (L * (L + 1)) / 2 : 1;
    P->heap_entries = 2u << P->po_lim;
    const size_t padn = (size_t)PADI(bs) + 1;
    P->samples_in_smem = (2 * padn * 4 <= 160 * 1024) ? 1 : 0;
    P->heap_in_smem = (P->heap_entries <= 2048) ? 1 : 0;
    P->samp_stride = (u32)padn;
    P->sr_code = sample_rate_code(p->sample_rate);
    P->bps_code = bps_code(p->bits_per_sample);
}

// ---- library ---------------------------------------------------------------
extern "C" int b200flac_abi_version(void) { return B200FLAC_ABI_VERSION; }

extern "C" int b200flac_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* b200flac_last_error(void) { return g_err; }

extern "C" void* b200flac_device_alloc(int device, uint64_t bytes)
{
    void* p = nullptr;
    CU_CHECK(cudaSetDevice(device), nullptr);
    CU_CHECK(cudaMalloc(&p, (size_t)bytes + 64), nullptr);
    return p;
}

extern "C" void b200flac_device_free(int device, void* ptr)
{
    if (cudaSetDevice(device) == cudaSuccess) cudaFree(ptr);
}

extern "C" int b200flac_device_upload(int device, void* dst, const void* src, uint64_t bytes)
{
    CU_CHECK(cudaSetDevice(device), 1);
    CU_CHECK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice), 1);
    return 0;
}

extern "C" int b200flac_device_download(int device, void* dst, const void* src, uint64_t bytes)
{
    CU_CHECK(cudaSetDevice(device), 1);
    CU_CHECK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost), 1);
    return 0;
}

extern "C" int b200flac_device_synth_pcm(int device, void* d_pcm, uint64_t seed, uint32_t channels,
                                         uint32_t bits_per_sample, uint64_t first_frame, uint64_t n_frames)
{
    CU_CHECK(cudaSetDevice(device), 1);
    k_synth_pcm<<<148 * 8, 256>>>((uint8_t*)d_pcm, seed, channels, bits_per_sample, first_frame, n_frames);
    CU_CHECK(cudaGetLastError(), 1);
    CU_CHECK(cudaDeviceSynchronize(), 1);
    return 0;
}

// ---- encoder ---------------------------------------------------------------
static u64 frame_bound_bytes(const b200flac_params* p, u32 n)
{
    // every subframe VERBATIM at bps+1 with a maximal wasted-bits field, plus header and CRC-16
    const u64 sub_bits = 8 + 32 + (u64)(p->bits_per_sample + 1) * n;
    return 16 + ((u64)p->channels * sub_bits + 7) / 8 + 2;
}

extern "C" uint64_t b200flac_encoder_output_bound(const b200flac_encoder* enc, uint64_t n_pcm_frames,
                                                  uint32_t n_segments)
{
    const b200flac_params* p = &enc->params;
    const u64 full = n_pcm_frames / p->block_size;
    u64 b = full * frame_bound_bytes(p, p->block_size) +
            ((u64)n_segments + 1) * frame_bound_bytes(p, p->block_size) + 64;
    if (p->no_verbatim_subframes) b *= 2; // FIXED/LPC are then not bounded by VERBATIM; see submit's guard
    return (b + 15) & ~15ull;
}

static void free_slot(Slot& s)
{
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.stream_b) { cudaStreamSynchronize(s.stream_b); cudaStreamDestroy(s.stream_b); }
    for (int i = 0; i < BF_MAX_MODEL_STREAMS; i++) {
        if (s.stream_hi[i]) { cudaStreamSynchronize(s.stream_hi[i]); cudaStreamDestroy(s.stream_hi[i]); }
        if (s.stream_lo[i]) { cudaStreamSynchronize(s.stream_lo[i]); cudaStreamDestroy(s.stream_lo[i]); }
    }
    for (int i = 0; i < BF_NUM_EVENTS; i++) if (s.ev[i]) cudaEventDestroy(s.ev[i]);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_in) cudaEventDestroy(s.ev_in);
    if (s.ev_join) cudaEventDestroy(s.ev_join);
    for (int i = 0; i < BF_MAX_CHUNKS; i++) {
        if (s.ev_model[i]) cudaEventDestroy(s.ev_model[i]);
        if (s.ev_an[i]) cudaEventDestroy(s.ev_an[i]);
        if (s.ev_scan[i]) cudaEventDestroy(s.ev_scan[i]);
    }
    cudaFree(s.d_chunk_tot);
    if (s.h_pcm) cudaFreeHost(s.h_pcm);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.h_fd) cudaFreeHost(s.h_fd);
    if (s.h_win) cudaFreeHost(s.h_win);
    if (s.h_frame_bytes) cudaFreeHost(s.h_frame_bytes);
    if (s.h_total) cudaFreeHost(s.h_total);
    cudaFree(s.d_pcm); cudaFree(s.d_fd); cudaFree(s.d_win); cudaFree(s.d_heads); cudaFree(s.d_coefs);
    cudaFree(s.d_plans); cudaFree(s.d_rice); cudaFree(s.d_choice); cudaFree(s.d_frame_bytes);
    cudaFree(s.d_frame_off); cudaFree(s.d_total); cudaFree(s.d_out); cudaFree(s.d_gsamples);
    cudaFree(s.d_gheap); cudaFree(s.d_gkarr);
    cudaFree(s.d_autoc); cudaFree(s.d_lpc_wasted); cudaFree(s.d_ticket); cudaFree(s.d_tasks);
    if (s.h_tasks) cudaFreeHost(s.h_tasks);
    if (s.h_odd) cudaFreeHost(s.h_odd);
    cudaFree(s.d_odd);
    if (s.stream) cudaStreamDestroy(s.stream);
}

extern "C" void b200flac_encoder_destroy(b200flac_encoder* enc)
{
    if (!enc) return;
    cudaSetDevice(enc->device);
    if (enc->slots) {
        for (int i = 0; i < enc->n_slots; i++) free_slot(enc->slots[i]);
        delete[] enc->slots;
    }
    cudaFree(enc->d_crc_tab);
    delete enc->windows;
    delete enc;
}

static u32 lpc_maxl(u32 L) { return L <= 8 ? 8 : L <= 12 ? 12 : L <= 16 ? 16 : 32; }

// staging mode of k_lpc_autoc for this stream shape (see the kernel)
static int lpc_mode(const bf_dev_params& P)
{
    if (P.stereo && P.bytes_ps == 2) return 0;
    return lpc_geometry(P.K, P.channels * P.bytes_ps, 13).bulk ? 1 : 2;
}

template <int MAXL, int MODE>
static cudaError_t lpc_prepare_one(b200flac_encoder* enc, size_t smem)
{
    // the opt-in limit is raised to the maximum so that a small batch can pad its request and
    // spread its few CTAs over all SMs (see launch_batch)
    const int lim = 200 * 1024;
    cudaError_t e = cudaFuncSetAttribute(k_lpc_autoc<MAXL, 2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&enc->lpc_occ[1], k_lpc_autoc<MAXL, 2, MODE>, 32, smem);
    if (e != cudaSuccess) return e;
    if constexpr (MAXL <= 16) {
        e = cudaFuncSetAttribute(k_lpc_autoc<MAXL, 1, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&enc->lpc_occ[0], k_lpc_autoc<MAXL, 1, MODE>, 32, smem);
    } else enc->lpc_occ[0] = 0;
    return e;
}

static cudaError_t lpc_prepare(b200flac_encoder* enc)
{
    const bf_dev_params& P = enc->P;
    cudaDeviceGetAttribute(&enc->n_sms, cudaDevAttrMultiProcessorCount, enc->device);
    const u32 maxl = lpc_maxl(P.max_lpc_order);
    const size_t smem = lpc_geometry(P.K, P.channels * P.bytes_ps, maxl + 1).smem_bytes;
#define LPC_PREP(MODE_) (maxl == 8 ? lpc_prepare_one<8, MODE_>(enc, smem) : maxl == 12 ? lpc_prepare_one<12, MODE_>(enc, smem) \
                         : maxl == 16 ? lpc_prepare_one<16, MODE_>(enc, smem) : lpc_prepare_one<32, MODE_>(enc, smem))
    const int mode = lpc_mode(P);
    return mode == 0 ? LPC_PREP(0) : mode == 1 ? LPC_PREP(1) : LPC_PREP(2);
#undef LPC_PREP
}

template <int S>
static cudaError_t set_smem_attrs(size_t smem_a, size_t smem_p)
{
    cudaError_t e = cudaFuncSetAttribute(k_analyze<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_pack_subframes<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
}

extern "C" b200flac_encoder* b200flac_encoder_create(const b200flac_params* params, int device,
                                                     uint64_t max_pcm_frames_per_batch, int n_slots)
{
    if (!params) { set_err("params is NULL"); return nullptr; }
    if (params->channels < 1 || params->channels > B200FLAC_MAX_CHANNELS) { set_err("unsupported channel count"); return nullptr; }
    if (params->bits_per_sample != 8 && params->bits_per_sample != 16 && params->bits_per_sample != 24) {
        set_err("bits_per_sample must be 8, 16 or 24"); return nullptr;
    }
    if (params->block_size < 1 || params->block_size > 65535) { set_err("block_size must be 1..65535"); return nullptr; }
    if (params->max_lpc_order > B200FLAC_MAX_LPC_ORDER) { set_err("max_lpc_order must be <= 32"); return nullptr; }
    if (n_slots < 1) n_slots = 1;
    if (max_pcm_frames_per_batch < params->block_size) max_pcm_frames_per_batch = params->block_size;
    int ndev = b200flac_device_count();
    if (ndev <= 0) { set_err("no CUDA device available: the B200 FLAC engine has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= ndev) { set_err("invalid CUDA device index"); return nullptr; }
    CU_CHECK(cudaSetDevice(device), nullptr);

    b200flac_encoder* enc = new b200flac_encoder();
    enc->params = *params;
    derive_params(params, &enc->P);
    enc->device = device;
    enc->max_pcm_frames = max_pcm_frames_per_batch;
    enc->n_slots = n_slots;
    enc->slots = nullptr;
    enc->windows = new std::map<u32, std::vector<double>>();
    enc->launches = 0;
    enc->chunk_frames = 0;     // one chunk: the pipeline measured no faster on B200
    enc->lookahead = 3;
    enc->model_streams = 4; enc->model_priority = 1; enc->lpc_grid_cap = 0;
    if (getenv("B200FLAC_NH")) enc->model_streams = (u32)atoi(getenv("B200FLAC_NH"));
    if (getenv("B200FLAC_LPC_PRIO")) enc->model_priority = (u32)atoi(getenv("B200FLAC_LPC_PRIO"));
    if (getenv("B200FLAC_LPC_GRID")) enc->lpc_grid_cap = (u32)atoi(getenv("B200FLAC_LPC_GRID"));
    if (getenv("B200FLAC_CHUNK")) enc->chunk_frames = (u32)atoi(getenv("B200FLAC_CHUNK"));        // tuning knobs
    if (getenv("B200FLAC_LOOKAHEAD")) enc->lookahead = (u32)std::max(1, atoi(getenv("B200FLAC_LOOKAHEAD")));
    const bf_dev_params& P = enc->P;

    const u32 bs = params->block_size;
    const size_t padn = (size_t)PADI(bs) + 1;
    // v2 kernels: whole block in one pass of <= 512 threads x S samples (S a multiple of 8), samples,
    // residual and partition heap in shared memory, VERBATIM a candidate (bounds the subframe image).
    enc->fast = false;
    enc->stage_words = 0;
    {
        int bestS = 0, bestNT = 0;
        long bestCost = -1;
        for (int cs = 8; cs <= 128; cs += 8) {
            int nt = (int)((bs + cs - 1) / cs);
            nt = (nt + 31) & ~31;
            if (nt > 512) continue;
            if (cs > 32 && bestS) break;                       // longer runs only when the block needs them
            const long cost = (long)nt * (cs + 24);            // per-thread fixed work ~ 24 samples' worth
            if (bestCost < 0 || cost < bestCost) { bestCost = cost; bestS = cs; bestNT = nt; }
        }
        const char* sforce = getenv("B200FLAC_S"); // tuning knob: force the samples-per-thread
        if (sforce) {
            const int fs = atoi(sforce);
            if (fs >= 8 && fs % 8 == 0) {
                int nt = (int)((bs + fs - 1) / fs);
                nt = (nt + 31) & ~31;
                if (nt <= 512) { bestS = fs; bestNT = nt; }
            }
        }
        const char* force = getenv("B200FLAC_FORCE_GENERIC");
        if (bestS && P.samples_in_smem && P.heap_in_smem && P.try_verbatim && !(force && force[0] == '1')) {
            enc->fast = true;
            enc->S = bestS;
            enc->NT = bestNT;
            const u64 sub_bits = 8 + 32 + (u64)(params->bits_per_sample + 1) * bs;
            enc->stage_words = (u32)((sub_bits + 31 + 31) / 32 + 2);
        }
    }
    size_t sa = 0, sp = 0;
    cudaError_t e = cudaSuccess;
    // k_analyze_v3: full-length blocks, every subframe type enabled, estimated LPC order, and a
    // finest partition that is a whole number of thread runs
    enc->v3 = false;
    {
        const char* force = getenv("B200FLAC_NO_V3");
        const u32 F = std::min<u32>(P.po_lim, (u32)__builtin_ctz(bs));
        if (enc->fast && P.try_lpc && P.try_fixed && P.try_constant && P.try_verbatim &&
            F <= V3_MAX_F + 1 && bs > P.max_lpc_order + 1 && !(force && force[0] == '1')) {
            enc->v3_sub = 1;
            if (F == V3_MAX_F + 1) {
                // partition order 8: 128 threads x 32 samples, a finest partition is half a thread run (run sums per half)
                if (bs == 32u * 128u && P.max_lpc_order <= 16) {
                    enc->v3 = true; enc->v3_S = 32; enc->v3_F = F; enc->v3_NT = 128; enc->v3_sub = 2;
                    enc->v3_smem = v3_smem_bytes(bs, 128, 2, P.exhaustive != 0);
                }
            } else
            for (u32 S3 = 32; S3 >= 8; S3 -= 8) {
                if (bs % S3 || (bs >> F) % S3) continue;
                const u32 nt = bs / S3;
                if (nt % 32 || nt < 32 || nt > 512) continue;
                const size_t sm = v3_smem_bytes(bs, nt, 1, P.exhaustive && S3 == 32 && nt == 128);
                if (sm > 200 * 1024) continue;
                enc->v3 = true; enc->v3_S = S3; enc->v3_F = F; enc->v3_NT = nt; enc->v3_smem = sm;
                break;
            }
        }
        if (enc->v3) {
            // samples per thread known at compile time (the exhaustive instantiation also assumes four warps)
            const bool s32 = enc->v3_S == 32 && (P.exhaustive ? enc->v3_NT == 128 : enc->v3_NT <= 128);
#define V3_ATTR(MINB_, EXH_, SC_) cudaFuncSetAttribute(k_analyze_v3<MINB_, EXH_, SC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN)
            // (orders up to 12 only: the instantiations without the 32-tap loops, for the shapes that have one)
            const bool short_lpc = P.max_lpc_order <= 12;
            if (enc->v3_sub == 2 && short_lpc)
                e = P.exhaustive ? cudaFuncSetAttribute(k_analyze_v3<5, true, 32, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN)
                                 : cudaFuncSetAttribute(k_analyze_v3<4, false, 32, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
            else if (enc->v3_sub == 2)
                e = P.exhaustive ? cudaFuncSetAttribute(k_analyze_v3<5, true, 32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN)
                                 : cudaFuncSetAttribute(k_analyze_v3<4, false, 32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
            else if (s32 && short_lpc)
                e = P.exhaustive ? cudaFuncSetAttribute(k_analyze_v3<5, true, 32, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN)
                                 : cudaFuncSetAttribute(k_analyze_v3<5, false, 32, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
            else if (P.exhaustive) e = s32 ? V3_ATTR(5, true, 32) : enc->v3_NT <= 128 ? V3_ATTR(5, true, 0) : enc->v3_NT <= 256 ? V3_ATTR(3, true, 0) : V3_ATTR(1, true, 0);
            else e = s32 ? V3_ATTR(5, false, 32) : enc->v3_NT <= 128 ? V3_ATTR(5, false, 0) : enc->v3_NT <= 256 ? V3_ATTR(3, false, 0) : V3_ATTR(1, false, 0);
#undef V3_ATTR
            if (e != cudaSuccess) enc->v3 = false;
            enc->v3_occ = 1;
            if (enc->v3) {
                if (enc->v3_sub == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&enc->v3_occ, k_analyze_v3<4, false, 32, 2>, (int)enc->v3_NT, enc->v3_smem);
                else if (enc->v3_NT <= 128) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&enc->v3_occ, k_analyze_v3<5, false, 0>, (int)enc->v3_NT, enc->v3_smem);
                else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&enc->v3_occ, k_analyze_v3<1, false, 0>, (int)enc->v3_NT, enc->v3_smem);
                if (enc->v3_occ < 1) enc->v3_occ = 1;
            }
            cudaDeviceGetAttribute(&enc->n_sms, cudaDevAttrMultiProcessorCount, enc->device);
            e = cudaSuccess;
        }
    }
    if (enc->fast) {
        sa = 2 * padn * 4 + 8 + (size_t)P.heap_entries * 9 + 2 * (size_t)P.rice_stride + 16;
        sp = 2 * padn * 4 + (size_t)enc->stage_words * 4 + 16;
        if (sa > 200 * 1024 || sp > 200 * 1024) enc->fast = false;
    }
    enc->p3 = false;
    enc->d_crc_tab = nullptr;
    if (enc->fast && !(getenv("B200FLAC_NO_P3") && getenv("B200FLAC_NO_P3")[0] == '1')) {
        // largest frame: every subframe VERBATIM, at most one of them (the side channel) at bps + 1
        const u64 fb = 16 + ((u64)params->channels * (8 + 32 + (u64)params->bits_per_sample * bs) + bs + 7) / 8 + 2;
        const u32 iw = (u32)((fb + 3) / 4 + 2);
        const size_t sm = p3_smem_bytes(bs, iw);
        // (a frame image that leaves room for a single CTA per SM -- six 24-bit channels -- packs faster
        // through the per-subframe kernels: 2.3 against 2.7 ms per 5 minutes of 96 kHz 5.1)
        if (sm <= 100 * 1024 && 2 * enc->NT <= 1024) {
            enc->p3 = true; enc->p3_img_words = iw; enc->p3_smem = sm;
            // CRC-16 tables: byte tables, x^(8 r) for r <= CHUNK, x^(8 * CHUNK * j)
            const u32 nchunks = (u32)(fb / P3_CHUNK_BYTES + 4);
            std::vector<unsigned short> t(1024 + P3_CHUNK_BYTES + 1 + nchunks);
            for (u32 b = 0; b < 256; b++) {
                u32 c = b << 8;
                for (int k = 0; k < 8; k++) c = (c & 0x8000) ? ((c << 1) ^ 0x8005) & 0xFFFF : (c << 1) & 0xFFFF;
                t[b] = (unsigned short)c;
            }
            // t[256 n + x]: CRC of byte x followed by n zero bytes (slicing by four)
            for (u32 n = 1; n < 4; n++)
                for (u32 b = 0; b < 256; b++) {
                    const u32 c = t[256 * (n - 1) + b];
                    t[256 * n + b] = (unsigned short)(((c << 8) & 0xFFFF) ^ t[c >> 8]);
                }
            auto mul = [](u32 a, u32 b) {
                u32 r = 0;
                for (int i = 0; i < 16; i++) if ((b >> i) & 1u) r ^= a << i;
                for (int i = 30; i >= 16; i--) if ((r >> i) & 1u) r ^= 0x18005u << (i - 16);
                return r;
            };
            u32 x8 = 0x100, acc = 1;                 // x^8
            for (u32 r = 0; r <= P3_CHUNK_BYTES; r++) { t[1024 + r] = (unsigned short)acc; acc = mul(acc, x8); }
            const u32 xc = t[1024 + P3_CHUNK_BYTES];  // x^(8 * CHUNK)
            acc = 1;
            for (u32 j = 0; j < nchunks; j++) { t[1024 + P3_CHUNK_BYTES + 1 + j] = (unsigned short)acc; acc = mul(acc, xc); }
            if (cudaMalloc((void**)&enc->d_crc_tab, t.size() * sizeof(unsigned short)) != cudaSuccess ||
                cudaMemcpy(enc->d_crc_tab, t.data(), t.size() * sizeof(unsigned short), cudaMemcpyHostToDevice) != cudaSuccess) {
                enc->p3 = false;
            } else {
                e = 2 * enc->NT <= 256 && enc->S == 32
                        ? cudaFuncSetAttribute(k_pack_v3<256, 4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN)
                    : 2 * enc->NT <= 256
                        ? cudaFuncSetAttribute(k_pack_v3<256, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN)
                        : cudaFuncSetAttribute(k_pack_v3<1024, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
                if (e != cudaSuccess) enc->p3 = false;
                e = cudaSuccess;
            }
        }
    }
    if (enc->fast) {
        enc->smem_analyze = sa;
        enc->smem_pack = sp;
        // two register budgets: small CTAs (<= 192 threads) are compiled for 6 resident CTAs per SM
        if (enc->NT <= 192) {
            e = cudaFuncSetAttribute(k_analyze_v2<192, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pack_v2<192, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
        } else {
            e = cudaFuncSetAttribute(k_analyze_v2<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pack_v2<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_OPTIN);
        }
    } else {
        enc->S = bs >= 2048 ? 32 : (bs >= 512 ? 16 : 8);
        int nt = (int)((bs + enc->S - 1) / enc->S);
        nt = (nt + 31) & ~31;
        if (nt < 32) nt = 32;
        if (nt > 512) nt = 512; // register budget: 512 threads x <=128 regs; longer blocks take several passes
        enc->NT = nt;
        sa = sp = 0;
        if (P.samples_in_smem) { sa += 2 * padn * 4; sp += padn * 4; }
        if (P.heap_in_smem) sa += 8 + (size_t)P.heap_entries * 9 + 2 * (size_t)P.rice_stride;
        enc->smem_analyze = sa + 16;
        enc->smem_pack = sp + 16;
        e = enc->S == 32 ? set_smem_attrs<32>(enc->smem_analyze, enc->smem_pack)
          : enc->S == 16 ? set_smem_attrs<16>(enc->smem_analyze, enc->smem_pack)
                         : set_smem_attrs<8>(enc->smem_analyze, enc->smem_pack);
    }
    if (e == cudaSuccess && P.try_lpc) e = lpc_prepare(enc);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        b200flac_encoder_destroy(enc);
        return nullptr;
    }

    const u64 M = max_pcm_frames_per_batch;
    const u64 maxf = M / bs + 2 + 1024; // room for up to 1024 segment tails per batch (checked in submit)
    enc->max_frames = (u32)maxf;
    const u64 U = maxf * P.K;
    enc->out_cap = M / bs * frame_bound_bytes(params, bs) + 1026 * frame_bound_bytes(params, bs) + 64;
    if (params->no_verbatim_subframes) enc->out_cap *= 2;
    enc->out_cap = (enc->out_cap + 15) & ~15ull;

    enc->slots = new Slot[n_slots];
    for (int i = 0; i < n_slots; i++) {
        Slot& s = enc->slots[i];
#define ALLOC(ptr, bytes) do { cudaError_t e2 = cudaMalloc((void**)&(ptr), (bytes)); if (e2 != cudaSuccess) { \
        snprintf(g_err, sizeof(g_err), "cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e2)); \
        b200flac_encoder_destroy(enc); return nullptr; } } while (0)
#define ALLOCH(ptr, bytes) do { cudaError_t e2 = cudaMallocHost((void**)&(ptr), (bytes)); if (e2 != cudaSuccess) { \
        snprintf(g_err, sizeof(g_err), "cudaMallocHost(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e2)); \
        b200flac_encoder_destroy(enc); return nullptr; } } while (0)
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
            set_err("cudaStreamCreate failed"); b200flac_encoder_destroy(enc); return nullptr;
        }
        for (int k = 0; k < BF_NUM_EVENTS; k++) cudaEventCreate(&s.ev[k]);
        cudaEventCreateWithFlags(&s.ev_done, cudaEventBlockingSync | cudaEventDisableTiming);
        ALLOCH(s.h_fd, maxf * sizeof(bf_frame_desc));
        ALLOCH(s.h_odd, maxf * sizeof(u32));
        ALLOC(s.d_odd, maxf * sizeof(u32));
        ALLOC(s.d_fd, maxf * sizeof(bf_frame_desc));
        s.win_cap = (size_t)bs * 4;
        ALLOCH(s.h_win, s.win_cap * sizeof(double));
        ALLOC(s.d_win, s.win_cap * sizeof(double) + 64);   // (+ slack: bulk copies round a tile up to 16 bytes)
        ALLOC(s.d_heads, U * sizeof(bf_lpc_head));
        ALLOC(s.d_coefs, U * P.model_stride * sizeof(short));
        if (P.try_lpc) {
            ALLOC(s.d_autoc, U * (size_t)(lpc_maxl(P.max_lpc_order) + 1) * sizeof(double));
            ALLOC(s.d_lpc_wasted, U);
            ALLOC(s.d_ticket, BF_MAX_CHUNKS * sizeof(u32));
            ALLOCH(s.h_tasks, maxf * sizeof(bf_lpc_task));
            ALLOC(s.d_tasks, maxf * sizeof(bf_lpc_task));
        }
        ALLOC(s.d_plans, U * sizeof(b200flac_plan));
        ALLOC(s.d_rice, U * P.rice_stride);
        ALLOC(s.d_choice, maxf * sizeof(bf_frame_choice));
        ALLOC(s.d_frame_bytes, maxf * sizeof(u32));
        ALLOC(s.d_frame_off, maxf * sizeof(u64));
        ALLOC(s.d_total, 64);
        ALLOC(s.d_chunk_tot, BF_MAX_CHUNKS * sizeof(u64));
        ALLOCH(s.h_frame_bytes, maxf * sizeof(u32));
        ALLOCH(s.h_total, 64);
        if (!P.samples_in_smem) ALLOC(s.d_gsamples, U * 2 * (size_t)P.samp_stride * sizeof(int));
        if (!P.heap_in_smem) {
            ALLOC(s.d_gheap, U * (size_t)P.heap_entries * sizeof(u64));
            ALLOC(s.d_gkarr, U * ((size_t)P.heap_entries + 2 * (size_t)P.rice_stride));
        }
#undef ALLOC
#undef ALLOCH
    }
    return enc;
}

// host-path buffers (device PCM + device output, pinned staging) are only needed by
// submit/collect; encode_device works on caller-provided device memory
static int ensure_host_path(b200flac_encoder* enc, Slot& s, bool want_staging)
{
    const size_t pcm_bytes = (size_t)enc->max_pcm_frames * enc->params.channels * enc->P.bytes_ps + 64;
    if (!s.d_pcm) CU_CHECK(cudaMalloc((void**)&s.d_pcm, pcm_bytes), 1);
    if (!s.d_out) CU_CHECK(cudaMalloc((void**)&s.d_out, enc->out_cap + 64), 1);
    if (want_staging && !s.h_pcm) CU_CHECK(cudaMallocHost((void**)&s.h_pcm, pcm_bytes), 1);
    return 0;
}

static bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int b200flac_internal_is_pinned(const void* p) { return is_pinned(p) ? 1 : 0; }   // (stream layer)

extern "C" uint8_t* b200flac_encoder_slot_pcm(b200flac_encoder* enc, int slot)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) return nullptr;
    if (cudaSetDevice(enc->device) != cudaSuccess) return nullptr;
    if (ensure_host_path(enc, enc->slots[slot], true)) return nullptr;
    return enc->slots[slot].h_pcm;
}

extern "C" uint8_t* b200flac_encoder_slot_out(b200flac_encoder* enc, int slot, uint64_t* capacity)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) return nullptr;
    if (cudaSetDevice(enc->device) != cudaSuccess) return nullptr;
    Slot& s = enc->slots[slot];
    if (!s.h_out) CU_CHECK(cudaMallocHost((void**)&s.h_out, enc->out_cap + 64), nullptr);
    if (capacity) *capacity = enc->out_cap;
    return s.h_out;
}

extern "C" void* b200flac_host_alloc(uint64_t bytes)
{
    void* p = nullptr;
    CU_CHECK(cudaMallocHost(&p, (size_t)bytes + 64), nullptr);
    return p;
}

extern "C" void b200flac_host_free(void* p) { if (p) cudaFreeHost(p); }

static const std::vector<double>& get_window(b200flac_encoder* enc, u32 n)
{
    auto it = enc->windows->find(n);
    if (it != enc->windows->end()) return it->second;
    std::vector<double> w(n);
    tukey_window(n, w.data());
    return enc->windows->emplace(n, std::move(w)).first->second;
}

// builds frame descriptors + windows for the batch; returns number of frames or -1
static long build_batch(b200flac_encoder* enc, Slot& s, const b200flac_segment* segs, u32 nseg, u64* pcm_frames_needed)
{
    const u32 bs = enc->params.block_size;
    std::map<u32, u32> woff;     // length -> offset in doubles
    size_t wused = 0;
    u64 nf = 0, need = 0;
    s.frame_pcm.clear();
    const bool want_windows = enc->P.try_lpc != 0;
    u32 last_n = 0xFFFFFFFFu, last_woff = 0;
    for (u32 g = 0; g < nseg; g++) {
        const b200flac_segment& sg = segs[g];
        if (sg.pcm_frame_offset + sg.n_pcm_frames > need) need = sg.pcm_frame_offset + sg.n_pcm_frames;
        u64 pos = 0;
        u32 fn = sg.first_frame_number;
        while (pos < sg.n_pcm_frames) {
            const u32 n = (u32)((sg.n_pcm_frames - pos) < bs ? (sg.n_pcm_frames - pos) : bs);
            if (nf >= enc->max_frames) { set_err("batch has more frames than the encoder was created for"); return -1; }
            bf_frame_desc& d = s.h_fd[nf];
            d.pcm_off = sg.pcm_frame_offset + pos;
            d.nsamp = n;
            d.frame_number = fn++;
            d.pad = 0;
            d.window_off = 0;
            if (want_windows) {
                if (n == last_n) {
                    d.window_off = last_woff;            // same length as the previous frame: no lookup
                } else {
                auto it = woff.find(n);
                if (it == woff.end()) {
                    if (wused + n > s.win_cap) {
                        // grow the window staging (rare: many distinct tail lengths in one batch)
                        size_t ncap = s.win_cap * 2 + n;
                        double *nh = nullptr, *nd = nullptr;
                        if (cudaMallocHost((void**)&nh, ncap * sizeof(double)) != cudaSuccess ||
                            cudaMalloc((void**)&nd, ncap * sizeof(double) + 64) != cudaSuccess) {
                            set_err("out of memory growing the window table"); return -1;
                        }
                        cudaStreamSynchronize(s.stream);
                        memcpy(nh, s.h_win, wused * sizeof(double));
                        cudaFreeHost(s.h_win); cudaFree(s.d_win);
                        s.h_win = nh; s.d_win = nd; s.win_cap = ncap;
                    }
                    const std::vector<double>& w = get_window(enc, n);
                    memcpy(s.h_win + wused, w.data(), (size_t)n * sizeof(double));
                    it = woff.emplace(n, (u32)wused).first;
                    wused += n;
                }
                d.window_off = it->second;
                last_n = n; last_woff = it->second;
                }
            }
            s.frame_pcm.push_back(n);
            nf++;
            pos += n;
        }
    }
    if (need > enc->max_pcm_frames) { set_err("batch has more PCM frames than the encoder was created for"); return -1; }
    // ---- chunks of the pipeline: frame ranges of about chunk_frames frames; every list below is cut at the
    // chunk boundaries and holds indices relative to its chunk's first frame ----
    s.chunks.clear();
    {
        u32 per_chunk = (u32)nf;
        if (enc->fast && enc->p3 && enc->chunk_frames && nf >= (u64)enc->chunk_frames + enc->chunk_frames / 2) {
            per_chunk = enc->chunk_frames;
            const u32 floor_ = (u32)((nf + BF_MAX_CHUNKS - 1) / BF_MAX_CHUNKS);
            if (per_chunk < floor_) per_chunk = floor_;
        }
        for (u64 f = 0; f < nf; f += per_chunk) {
            Chunk c;
            memset(&c, 0, sizeof(c));
            c.frame0 = (u32)f;
            c.n_frames = (u32)std::min<u64>(per_chunk, nf - f);
            // (a short last chunk joins its predecessor)
            if (nf - f - c.n_frames < per_chunk / 2) c.n_frames = (u32)(nf - f);
            s.chunks.push_back(c);
            if (c.n_frames != per_chunk) break;
        }
    }
    s.n_odd = 0;
    s.n_tasks = 0;
    for (Chunk& c : s.chunks) {
        c.odd0 = s.n_odd;
        for (u32 f = 0; f < c.n_frames; f++) if (s.h_fd[c.frame0 + f].nsamp != bs) s.h_odd[s.n_odd++] = f;
        c.n_odd = s.n_odd - c.odd0;
        c.task0 = s.n_tasks;
        if (want_windows) {
            // tasks of the autocorrelation kernel: runs of up to 32/K consecutive frames of one length
            const u32 per = 32 / enc->P.K;
            const bf_frame_desc* fd = s.h_fd + c.frame0;
            for (u32 f = 0; f < c.n_frames;) {
                u32 k = 1;
                while (k < per && f + k < c.n_frames && fd[f + k].nsamp == fd[f].nsamp) k++;
                s.h_tasks[s.n_tasks].first_frame = f;
                s.h_tasks[s.n_tasks].n_frames = k;
                s.n_tasks++;
                f += k;
            }
        }
        c.n_tasks = s.n_tasks - c.task0;
    }
    *pcm_frames_needed = need;
    s.n_frames = (u32)nf;
    // windows upload size is remembered through woff/wused: stash in h_total[1]
    s.h_total[1] = wused;
    return (long)nf;
}

template <int S>
static void launch_analyze_pack(b200flac_encoder* enc, Slot& s, const uint8_t* d_pcm, uint8_t* d_out, u64 out_cap)
{
    const bf_dev_params& P = enc->P;
    const u32 nf = s.n_frames, U = nf * P.K;
    cudaStream_t st = s.stream;
    k_analyze<S><<<U, enc->NT, enc->smem_analyze, st>>>(d_pcm, s.d_fd, P, s.d_heads, s.d_coefs, s.d_plans, s.d_rice,
                                                       s.d_gsamples, s.d_gheap, s.d_gkarr);
    cudaEventRecord(s.ev[2], st);
    k_frame_select<<<(nf + 127) / 128, 128, 0, st>>>(s.d_fd, nf, P, s.d_plans, s.d_choice, s.d_frame_bytes);
    k_scan_offsets<<<1, 1024, 0, st>>>(s.d_frame_bytes, nf, s.d_frame_off, s.d_total);
    k_zero_output<<<148 * 4, 256, 0, st>>>((uint4*)d_out, s.d_total, out_cap);
    cudaEventRecord(s.ev[3], st);
    k_pack_subframes<S><<<nf * P.channels, enc->NT, enc->smem_pack, st>>>(d_pcm, s.d_fd, P, s.d_plans, s.d_rice,
                                                                         s.d_choice, s.d_frame_off, (u32*)d_out,
                                                                         s.d_gsamples, s.d_total, out_cap);
    cudaEventRecord(s.ev[4], st);
    k_frame_crc16<<<(nf * 32 + 127) / 128, 128, 0, st>>>(s.d_frame_off, s.d_frame_bytes, nf, d_out, s.d_total, out_cap);
    cudaEventRecord(s.ev[5], st);
    count_launches(enc, 6);
}

// the slot's arrays as one chunk sees them: from its first frame / first unit
struct ChunkView {
    const bf_frame_desc* fd;
    u32 nf, U;
    bf_lpc_head* heads;
    short* coefs;
    double* autoc;
    uint8_t* wasted;
    b200flac_plan* plans;
    uint8_t* rice;
    bf_frame_choice* choice;
    u32* frame_bytes;
    u64* frame_off;
    const bf_lpc_task* tasks;
    u32 n_tasks;
    const u32* odd;
    u32 n_odd;
    u32* ticket;
};

static ChunkView chunk_view(const b200flac_encoder* enc, const Slot& s, u32 ci)
{
    const bf_dev_params& P = enc->P;
    const Chunk& c = s.chunks[ci];
    const size_t u0 = (size_t)c.frame0 * P.K;
    ChunkView v;
    v.fd = s.d_fd + c.frame0;
    v.nf = c.n_frames;
    v.U = c.n_frames * P.K;
    v.heads = s.d_heads + u0;
    v.coefs = s.d_coefs + u0 * P.model_stride;
    v.autoc = s.d_autoc ? s.d_autoc + u0 * (lpc_maxl(P.max_lpc_order) + 1) : nullptr;
    v.wasted = s.d_lpc_wasted ? s.d_lpc_wasted + u0 : nullptr;
    v.plans = s.d_plans + u0;
    v.rice = s.d_rice + u0 * P.rice_stride;
    v.choice = s.d_choice + c.frame0;
    v.frame_bytes = s.d_frame_bytes + c.frame0;
    v.frame_off = s.d_frame_off + c.frame0;
    v.tasks = s.d_tasks ? s.d_tasks + c.task0 : nullptr;
    v.n_tasks = c.n_tasks;
    v.odd = s.d_odd + c.odd0;
    v.n_odd = c.n_odd;
    v.ticket = s.d_ticket ? s.d_ticket + ci : nullptr;
    return v;
}

// window + autocorrelation, then Levinson-Durbin / order estimate / quantisation, of one chunk
static void stage_model(b200flac_encoder* enc, const ChunkView& v, cudaStream_t st, const uint8_t* d_pcm,
                        const double* d_win, bool piped)
{
    const bf_dev_params& P = enc->P;
    // persistent one-warp CTAs drawing (32 units, lag group) tasks from a ticket counter; the lag
    // split halves a thread's sequential work, which is what a small batch's duration is made of --
    // inside the pipeline the other chunks' kernels fill the machine, so the split's duplicated unpacking
    // is not paid there
    const u32 maxl = lpc_maxl(P.max_lpc_order);
    const u32 groups = v.n_tasks;
    bool split = maxl == 32 || (!piped && groups < (u32)enc->n_sms * 12u);
    if (getenv("B200FLAC_LPC_G") && maxl != 32) split = atoi(getenv("B200FLAC_LPC_G")) == 2; // tuning knob
    const u32 G = split ? 2 : 1;
    const u32 n_tasks = groups * G;
    const int occ = enc->lpc_occ[G - 1] > 0 ? enc->lpc_occ[G - 1] : 1;
    size_t lsm = lpc_geometry(P.K, P.channels * P.bytes_ps, maxl + 1).smem_bytes;
    u32 grid = (u32)enc->n_sms * (u32)occ;
    if (piped && enc->lpc_grid_cap && grid > enc->lpc_grid_cap) grid = enc->lpc_grid_cap;
    if (n_tasks < grid) {
        grid = n_tasks;
        if (!piped) {
            // few tasks: the block scheduler fills an SM before it moves to the next, so pad the
            // shared-memory request until only ceil(tasks / SMs) CTAs fit on one
            const u32 per_sm = (n_tasks + enc->n_sms - 1) / enc->n_sms;
            const size_t pad = (size_t)(220 * 1024) / per_sm - 1024;
            if (pad > lsm) lsm = pad < (size_t)(200 * 1024) ? pad : (size_t)(200 * 1024);
        }
    }
    const int mode = lpc_mode(P);
#define LPC_LAUNCH(MAXL_, G_) do { \
        if (mode == 0) k_lpc_autoc<MAXL_, G_, 0><<<grid, 32, lsm, st>>>(d_pcm, v.fd, v.nf, d_win, P, v.tasks, n_tasks, v.ticket, v.autoc, v.wasted); \
        else if (mode == 1) k_lpc_autoc<MAXL_, G_, 1><<<grid, 32, lsm, st>>>(d_pcm, v.fd, v.nf, d_win, P, v.tasks, n_tasks, v.ticket, v.autoc, v.wasted); \
        else k_lpc_autoc<MAXL_, G_, 2><<<grid, 32, lsm, st>>>(d_pcm, v.fd, v.nf, d_win, P, v.tasks, n_tasks, v.ticket, v.autoc, v.wasted); \
    } while (0)
    if (maxl == 8) { if (split) LPC_LAUNCH(8, 2); else LPC_LAUNCH(8, 1); }
    else if (maxl == 12) { if (split) LPC_LAUNCH(12, 2); else LPC_LAUNCH(12, 1); }
    else if (maxl == 16) { if (split) LPC_LAUNCH(16, 2); else LPC_LAUNCH(16, 1); }
    else LPC_LAUNCH(32, 2);
#undef LPC_LAUNCH
    const u32 fb = (v.U + 127) / 128;
    if (maxl == 8) k_lpc_finish<8><<<fb, 128, 0, st>>>(v.fd, v.nf, P, v.autoc, v.wasted, v.heads, v.coefs);
    else if (maxl == 12) k_lpc_finish<12><<<fb, 128, 0, st>>>(v.fd, v.nf, P, v.autoc, v.wasted, v.heads, v.coefs);
    else if (maxl == 16) k_lpc_finish<16><<<fb, 128, 0, st>>>(v.fd, v.nf, P, v.autoc, v.wasted, v.heads, v.coefs);
    else k_lpc_finish<32><<<fb, 128, 0, st>>>(v.fd, v.nf, P, v.autoc, v.wasted, v.heads, v.coefs);
    count_launches(enc, 2);
}

// model search of one chunk (fast kernels)
static void stage_analyze(b200flac_encoder* enc, const ChunkView& v, cudaStream_t st, const uint8_t* d_pcm)
{
    const bf_dev_params& P = enc->P;
    u32 gridv2 = v.U;
    const u32* flist = nullptr;
    if (enc->v3) {
        // one CTA per unit.  (The kernel also runs as a persistent grid -- B200FLAC_V3_GRID=740 is one wave
        // -- but CTAs that start together stay in the same phase of the unit and overlap their load and
        // search phases worse: 3.16 ms against 2.87 ms per hour on B200.)
        u32 g3 = v.U;
        if (getenv("B200FLAC_V3_GRID")) g3 = (u32)atoi(getenv("B200FLAC_V3_GRID"));   // tuning knob
        if (g3 > v.U || g3 == 0) g3 = v.U;
#define V3_LAUNCH(MINB_, EXH_, SC_) k_analyze_v3<MINB_, EXH_, SC_><<<g3, enc->v3_NT, enc->v3_smem, st>>>(d_pcm, v.fd, P, enc->v3_S, enc->v3_F, v.U, v.heads, v.coefs, v.plans, v.rice)
#define V3_LAUNCH5(MINB_, EXH_, SUB_, LONG_) k_analyze_v3<MINB_, EXH_, 32, SUB_, LONG_><<<g3, enc->v3_NT, enc->v3_smem, st>>>(d_pcm, v.fd, P, enc->v3_S, enc->v3_F, v.U, v.heads, v.coefs, v.plans, v.rice)
        const bool short_lpc = P.max_lpc_order <= 12;
        if (enc->v3_sub == 2) {
            if (P.exhaustive) { if (short_lpc) V3_LAUNCH5(5, true, 2, false); else V3_LAUNCH5(5, true, 2, true); }
            else { if (short_lpc) V3_LAUNCH5(4, false, 2, false); else V3_LAUNCH5(4, false, 2, true); }
        }
        else if (enc->v3_S == 32 && (P.exhaustive ? enc->v3_NT == 128 : enc->v3_NT <= 128)) {
            if (P.exhaustive) { if (short_lpc) V3_LAUNCH5(5, true, 1, false); else V3_LAUNCH(5, true, 32); }
            else { if (short_lpc) V3_LAUNCH5(5, false, 1, false); else V3_LAUNCH(5, false, 32); }
        }
#undef V3_LAUNCH5
        else if (enc->v3_NT <= 128) { if (P.exhaustive) V3_LAUNCH(5, true, 0); else V3_LAUNCH(5, false, 0); }
        else if (enc->v3_NT <= 256) { if (P.exhaustive) V3_LAUNCH(3, true, 0); else V3_LAUNCH(3, false, 0); }
        else { if (P.exhaustive) V3_LAUNCH(1, true, 0); else V3_LAUNCH(1, false, 0); }
#undef V3_LAUNCH
        count_launches(enc, 1);
        gridv2 = v.n_odd * P.K;      // the other block lengths (a stream's last block)
        flist = v.odd;
    }
    if (gridv2) {
        if (enc->NT <= 192)
            k_analyze_v2<192, 4><<<gridv2, enc->NT, enc->smem_analyze, st>>>(d_pcm, v.fd, P, (u32)enc->S, v.heads, v.coefs, v.plans, v.rice, flist);
        else
            k_analyze_v2<512, 1><<<gridv2, enc->NT, enc->smem_analyze, st>>>(d_pcm, v.fd, P, (u32)enc->S, v.heads, v.coefs, v.plans, v.rice, flist);
        count_launches(enc, 1);
    }
}

static void stage_pack_v3(b200flac_encoder* enc, const ChunkView& v, cudaStream_t st, const uint8_t* d_pcm,
                          uint8_t* d_out, u64 out_cap, const u64* d_total)
{
    const bf_dev_params& P = enc->P;
#define P3_LAUNCH(NTMAX_, MINB_, SC_) k_pack_v3<NTMAX_, MINB_, SC_><<<v.nf, 2 * enc->NT, enc->p3_smem, st>>>( \
            d_pcm, v.fd, P, (u32)enc->S, v.plans, v.rice, v.choice, v.frame_off, d_out, d_total, out_cap, \
            enc->p3_img_words, enc->d_crc_tab, enc->d_crc_tab + 1024)
    if (2 * enc->NT <= 256 && enc->S == 32) P3_LAUNCH(256, 4, 32);
    else if (2 * enc->NT <= 256) P3_LAUNCH(256, 4, 0);
    else P3_LAUNCH(1024, 1, 0);
#undef P3_LAUNCH
    count_launches(enc, 1);
}

// the fast kernels of a batch that is a single chunk: back to back on the slot's stream, one event between
// stages (b200flac_encoder_last_kernel_ms)
static void launch_analyze_pack_v2(b200flac_encoder* enc, Slot& s, const uint8_t* d_pcm, uint8_t* d_out, u64 out_cap)
{
    const bf_dev_params& P = enc->P;
    const ChunkView v = chunk_view(enc, s, 0);
    const u32 nf = v.nf;
    cudaStream_t st = s.stream;
    stage_analyze(enc, v, st, d_pcm);
    cudaEventRecord(s.ev[2], st);
    k_frame_select<<<(nf + 127) / 128, 128, 0, st>>>(s.d_fd, nf, P, s.d_plans, s.d_choice, s.d_frame_bytes);
    k_scan_offsets<<<1, 1024, 0, st>>>(s.d_frame_bytes, nf, s.d_frame_off, s.d_total);
    count_launches(enc, 2);
    if (enc->p3) {
        cudaEventRecord(s.ev[3], st);
        stage_pack_v3(enc, v, st, d_pcm, d_out, out_cap, s.d_total);
        cudaEventRecord(s.ev[4], st);
        cudaEventRecord(s.ev[5], st);
        return;
    }
    k_zero_output<<<148 * 4, 256, 0, st>>>((uint4*)d_out, s.d_total, out_cap);
    cudaEventRecord(s.ev[3], st);
    if (enc->NT <= 192)
        k_pack_v2<192, 4><<<nf * P.channels, enc->NT, enc->smem_pack, st>>>(d_pcm, s.d_fd, P, (u32)enc->S, s.d_plans, s.d_rice,
                                                                         s.d_choice, s.d_frame_off, (u32*)d_out, s.d_total,
                                                                         out_cap, enc->stage_words);
    else
        k_pack_v2<512, 1><<<nf * P.channels, enc->NT, enc->smem_pack, st>>>(d_pcm, s.d_fd, P, (u32)enc->S, s.d_plans, s.d_rice,
                                                                         s.d_choice, s.d_frame_off, (u32*)d_out, s.d_total,
                                                                         out_cap, enc->stage_words);
    cudaEventRecord(s.ev[4], st);
    k_frame_crc16<<<(nf * 32 + 127) / 128, 128, 0, st>>>(s.d_frame_off, s.d_frame_bytes, nf, d_out, s.d_total, out_cap);
    cudaEventRecord(s.ev[5], st);
    count_launches(enc, 3);
}

// streams and events of the chunk pipeline, created on a slot's first pipelined batch
static int ensure_pipeline(Slot& s)
{
    if (s.stream_b) return 0;
    int lo = 0, hi = 0;
    CU_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi), 1);
    CU_CHECK(cudaStreamCreateWithFlags(&s.stream_b, cudaStreamNonBlocking), 1);
    for (int i = 0; i < BF_MAX_MODEL_STREAMS; i++) {
        CU_CHECK(cudaStreamCreateWithPriority(&s.stream_hi[i], cudaStreamNonBlocking, hi), 1);
        CU_CHECK(cudaStreamCreateWithFlags(&s.stream_lo[i], cudaStreamNonBlocking), 1);
    }
    CU_CHECK(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming), 1);
    CU_CHECK(cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming), 1);
    for (int i = 0; i < BF_MAX_CHUNKS; i++) {
        CU_CHECK(cudaEventCreateWithFlags(&s.ev_model[i], cudaEventDisableTiming), 1);
        CU_CHECK(cudaEventCreateWithFlags(&s.ev_an[i], cudaEventDisableTiming), 1);
        CU_CHECK(cudaEventCreateWithFlags(&s.ev_scan[i], cudaEventDisableTiming), 1);
    }
    return 0;
}

// A batch of several chunks as a software pipeline over three streams.  The floating-point model kernels
// (FP64 pipe) of chunk c + 1 .. c + lookahead run on a high-priority stream while the integer kernels (ALU /
// FMA pipes) of chunk c run on two alternating streams, so the pipes that sat idle in turn are busy together,
// and a chunk's PCM (~33 MB) is still in the 126 MB L2 when its second and third readers come.  The frame
// offsets are chained through the per-chunk running totals (k_scan_offsets' carry-in).
static int launch_pipeline(b200flac_encoder* enc, Slot& s, const uint8_t* d_pcm, uint8_t* d_out, u64 out_cap)
{
    const bf_dev_params& P = enc->P;
    if (ensure_pipeline(s)) return 1;
    const u32 NC = (u32)s.chunks.size();
    cudaStream_t main_st = s.stream;
    u32 NH = enc->model_streams;
    if (NH < 1) NH = 1;
    if (NH > BF_MAX_MODEL_STREAMS) NH = BF_MAX_MODEL_STREAMS;
    cudaStream_t* hs = enc->model_priority ? s.stream_hi : s.stream_lo;
    const u32 la = std::max(enc->lookahead, NH);
    if (P.try_lpc) cudaMemsetAsync(s.d_ticket, 0, NC * sizeof(u32), main_st);
    cudaEventRecord(s.ev_in, main_st);                 // inputs (PCM, descriptors, tasks, windows) are on the device
    cudaStreamWaitEvent(s.stream_b, s.ev_in, 0);
    if (P.try_lpc) for (u32 h = 0; h < NH; h++) cudaStreamWaitEvent(hs[h], s.ev_in, 0);
    for (u32 ci = 0; ci < NC; ci++) {
        const ChunkView v = chunk_view(enc, s, ci);
        cudaStream_t st = (ci & 1) ? s.stream_b : main_st;
        if (P.try_lpc) {
            cudaStream_t h = hs[ci % NH];
            if (ci >= la) cudaStreamWaitEvent(h, s.ev_an[ci - la], 0);
            stage_model(enc, v, h, d_pcm, s.d_win, true);
            cudaEventRecord(s.ev_model[ci], h);
            cudaStreamWaitEvent(st, s.ev_model[ci], 0);
        }
        stage_analyze(enc, v, st, d_pcm);
        cudaEventRecord(s.ev_an[ci], st);
        k_frame_select<<<(v.nf + 127) / 128, 128, 0, st>>>(v.fd, v.nf, P, v.plans, v.choice, v.frame_bytes);
        if (ci) cudaStreamWaitEvent(st, s.ev_scan[ci - 1], 0);
        k_scan_offsets<<<1, 1024, 0, st>>>(v.frame_bytes, v.nf, v.frame_off, s.d_chunk_tot + ci,
                                           ci ? s.d_chunk_tot + ci - 1 : nullptr, ci + 1 == NC ? s.d_total : nullptr);
        cudaEventRecord(s.ev_scan[ci], st);
        stage_pack_v3(enc, v, st, d_pcm, d_out, out_cap, s.d_chunk_tot + ci);
        count_launches(enc, 2);
    }
    cudaEventRecord(s.ev_join, s.stream_b);
    cudaStreamWaitEvent(main_st, s.ev_join, 0);
    for (int i = 1; i <= 5; i++) cudaEventRecord(s.ev[i], main_st);
    return 0;
}

// enqueue the kernels of one batch
static int launch_batch(b200flac_encoder* enc, Slot& s, const uint8_t* d_pcm, uint8_t* d_out, u64 out_cap)
{
    const bf_dev_params& P = enc->P;
    cudaStream_t st = s.stream;
    cudaEventRecord(s.ev[0], st);
    if (s.chunks.size() > 1) {
        if (launch_pipeline(enc, s, d_pcm, d_out, out_cap)) return 1;
        CU_CHECK(cudaGetLastError(), 1);
        s.timed = false;          // kernels of different chunks overlap: there is no per-kernel time
        return 0;
    }
    if (P.try_lpc) {
        cudaMemsetAsync(s.d_ticket, 0, sizeof(u32), st);
        stage_model(enc, chunk_view(enc, s, 0), st, d_pcm, s.d_win, false);
    }
    cudaEventRecord(s.ev[1], st);
    if (enc->fast) {
        launch_analyze_pack_v2(enc, s, d_pcm, d_out, out_cap);
    } else if (enc->S == 32) launch_analyze_pack<32>(enc, s, d_pcm, d_out, out_cap);
    else if (enc->S == 16) launch_analyze_pack<16>(enc, s, d_pcm, d_out, out_cap);
    else launch_analyze_pack<8>(enc, s, d_pcm, d_out, out_cap);
    CU_CHECK(cudaGetLastError(), 1);
    s.timed = true;
    return 0;
}

extern "C" int b200flac_encoder_submit(b200flac_encoder* enc, int slot, const uint8_t* pcm,
                                       const b200flac_segment* segments, uint32_t n_segments)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) { set_err("bad encoder or slot"); return 1; }
    if (!pcm || !segments || n_segments == 0 || n_segments > BF_MAX_SEGMENTS) { set_err("bad segments"); return 1; }
    Slot& s = enc->slots[slot];
    if (s.busy) { set_err("slot is busy: collect it first"); return 1; }
    CU_CHECK(cudaSetDevice(enc->device), 1);
    u64 need = 0;
    const long nf = build_batch(enc, s, segments, n_segments, &need);
    if (nf < 0) return 1;
    if (nf == 0) { s.busy = true; s.timed = false; *s.h_total = 0; return 0; }
    const size_t pcm_bytes = (size_t)need * enc->params.channels * enc->P.bytes_ps;
    cudaStream_t st = s.stream;
    const uint8_t* src = pcm;
    if (pcm != s.h_pcm && !is_pinned(pcm)) {
        // pageable caller memory: stage through the slot's pinned buffer so the copy is asynchronous
        if (ensure_host_path(enc, s, true)) return 1;
        memcpy(s.h_pcm, pcm, pcm_bytes);
        src = s.h_pcm;
    } else if (ensure_host_path(enc, s, false)) return 1;
    CU_CHECK(cudaMemcpyAsync(s.d_pcm, src, pcm_bytes, cudaMemcpyHostToDevice, st), 1);
    CU_CHECK(cudaMemcpyAsync(s.d_fd, s.h_fd, (size_t)nf * sizeof(bf_frame_desc), cudaMemcpyHostToDevice, st), 1);
    if (s.n_odd) CU_CHECK(cudaMemcpyAsync(s.d_odd, s.h_odd, (size_t)s.n_odd * sizeof(u32), cudaMemcpyHostToDevice, st), 1);
    if (enc->P.try_lpc) CU_CHECK(cudaMemcpyAsync(s.d_tasks, s.h_tasks, (size_t)s.n_tasks * sizeof(bf_lpc_task), cudaMemcpyHostToDevice, st), 1);
    const size_t wused = (size_t)s.h_total[1];
    if (wused) CU_CHECK(cudaMemcpyAsync(s.d_win, s.h_win, wused * sizeof(double), cudaMemcpyHostToDevice, st), 1);
    if (launch_batch(enc, s, s.d_pcm, s.d_out, enc->out_cap)) return 1;
    CU_CHECK(cudaMemcpyAsync(s.h_frame_bytes, s.d_frame_bytes, (size_t)nf * sizeof(u32), cudaMemcpyDeviceToHost, st), 1);
    CU_CHECK(cudaMemcpyAsync(s.h_total, s.d_total, sizeof(u64), cudaMemcpyDeviceToHost, st), 1);
    CU_CHECK(cudaEventRecord(s.ev_done, st), 1);
    s.busy = true;
    return 0;
}

extern "C" int b200flac_encoder_collect(b200flac_encoder* enc, int slot, uint8_t* out, uint64_t out_capacity,
                                        uint64_t* out_bytes, uint32_t* frame_bytes, uint32_t* frame_pcm,
                                        uint32_t frame_capacity, uint32_t* n_frames)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) { set_err("bad encoder or slot"); return 1; }
    Slot& s = enc->slots[slot];
    if (!s.busy) { set_err("slot has no batch in flight"); return 1; }
    CU_CHECK(cudaSetDevice(enc->device), 1);
    s.busy = false;
    if (s.n_frames == 0) { if (out_bytes) *out_bytes = 0; if (n_frames) *n_frames = 0; return 0; }
    // (an event with cudaEventBlockingSync: the many-streams case runs one host thread per stream next to the
    // streams' MD5 threads, and a spinning wait took a core from them)
    CU_CHECK(wait_event(s.ev_done), 1);
    const u64 total = *s.h_total;
    if (total + 16 > enc->out_cap) { set_err("encoded batch exceeds the device output buffer (VERBATIM disabled?)"); return 1; }
    if (total > out_capacity) { set_err("output buffer too small"); return 1; }
    if (s.n_frames > frame_capacity && (frame_bytes || frame_pcm)) { set_err("frame arrays too small"); return 1; }
    CU_CHECK(cudaMemcpyAsync(out, s.d_out, (size_t)total, cudaMemcpyDeviceToHost, s.stream), 1);
    CU_CHECK(cudaEventRecord(s.ev_done, s.stream), 1);
    if (frame_bytes) memcpy(frame_bytes, s.h_frame_bytes, (size_t)s.n_frames * sizeof(u32));
    if (frame_pcm) memcpy(frame_pcm, s.frame_pcm.data(), (size_t)s.n_frames * sizeof(u32));
    CU_CHECK(wait_event(s.ev_done), 1);
    if (out_bytes) *out_bytes = total;
    if (n_frames) *n_frames = s.n_frames;
    return 0;
}

extern "C" int b200flac_encoder_encode(b200flac_encoder* enc, const uint8_t* pcm,
                                       const b200flac_segment* segments, uint32_t n_segments,
                                       uint8_t* out, uint64_t out_capacity, uint64_t* out_bytes,
                                       uint32_t* frame_bytes, uint32_t* frame_pcm, uint32_t frame_capacity,
                                       uint32_t* n_frames)
{
    if (b200flac_encoder_submit(enc, 0, pcm, segments, n_segments)) return 1;
    return b200flac_encoder_collect(enc, 0, out, out_capacity, out_bytes, frame_bytes, frame_pcm, frame_capacity, n_frames);
}

// Small transfers between page-locked host memory and the device done by a kernel over the mapped host pointer
// instead of a copy-engine job.  The device-resident entry is used by callers that keep the copy engines busy with
// their own traffic (b200flac_encode_files queues hundreds of MB of PCM ahead of the encoder): a batch's descriptors
// queued as an ordinary copy wait behind everything already in the engine's queue, and the kernels with them.
__global__ void k_copy_words(u32* __restrict__ dst, const u32* __restrict__ src, size_t n_words)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
static void copy_by_kernel(b200flac_encoder* enc, void* dst, const void* src, size_t bytes, cudaStream_t st)
{
    const size_t n = bytes / 4;       // (every table copied this way is a whole number of words)
    if (!n) return;
    count_launches(enc, 1);
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 296);
    k_copy_words<<<blocks, 256, 0, st>>>((u32*)dst, (const u32*)src, n);
}

// device-resident batches, asynchronous: enqueue (host work + launches) now, wait later -- with two slots the
// host's share of batch k + 1 (descriptors, task lists, their upload) runs under the kernels of batch k
extern "C" int b200flac_encoder_submit_device(b200flac_encoder* enc, int slot, const void* d_pcm,
                                              const b200flac_segment* segments, uint32_t n_segments,
                                              void* d_out, uint64_t out_capacity)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) { set_err("bad encoder or slot"); return 1; }
    if (!d_pcm || !d_out || !segments || n_segments == 0) { set_err("bad arguments"); return 1; }
    if (((uintptr_t)d_out & 15) || ((uintptr_t)d_pcm & 3)) { set_err("d_out must be 16-byte and d_pcm 4-byte aligned"); return 1; }
    Slot& s = enc->slots[slot];
    if (s.busy) { set_err("slot is busy"); return 1; }
    CU_CHECK(cudaSetDevice(enc->device), 1);
    u64 need = 0;
    const long nf = build_batch(enc, s, segments, n_segments, &need);
    if (nf < 0) return 1;
    s.dev_cap = out_capacity & ~15ull;
    s.busy = true;
    if (nf == 0) { s.timed = false; *s.h_total = 0; return 0; }
    cudaStream_t st = s.stream;
    copy_by_kernel(enc, s.d_fd, s.h_fd, (size_t)nf * sizeof(bf_frame_desc), st);
    if (s.n_odd) copy_by_kernel(enc, s.d_odd, s.h_odd, (size_t)s.n_odd * sizeof(u32), st);
    if (enc->P.try_lpc) copy_by_kernel(enc, s.d_tasks, s.h_tasks, (size_t)s.n_tasks * sizeof(bf_lpc_task), st);
    const size_t wused = (size_t)s.h_total[1];
    if (wused) copy_by_kernel(enc, s.d_win, s.h_win, wused * sizeof(double), st);
    if (launch_batch(enc, s, (const uint8_t*)d_pcm, (uint8_t*)d_out, s.dev_cap)) { s.busy = false; return 1; }
    copy_by_kernel(enc, s.h_total, s.d_total, sizeof(u64), st);
    copy_by_kernel(enc, s.h_frame_bytes, s.d_frame_bytes, (size_t)nf * sizeof(u32), st);
    CU_CHECK(cudaGetLastError(), 1);
    CU_CHECK(cudaEventRecord(s.ev_done, st), 1);
    return 0;
}

extern "C" int b200flac_encoder_collect_device(b200flac_encoder* enc, int slot, uint64_t* out_bytes,
                                               uint32_t* n_frames, float* elapsed_ms)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) { set_err("bad encoder or slot"); return 1; }
    Slot& s = enc->slots[slot];
    if (!s.busy) { set_err("slot has no batch in flight"); return 1; }
    CU_CHECK(cudaSetDevice(enc->device), 1);
    s.busy = false;
    if (s.n_frames == 0) { if (out_bytes) *out_bytes = 0; if (n_frames) *n_frames = 0; return 0; }
    CU_CHECK(wait_event(s.ev_done), 1);
    const u64 total = *s.h_total;
    if (total + 16 > s.dev_cap) { set_err("encoded batch exceeds the output buffer"); return 1; }
    if (out_bytes) *out_bytes = total;
    if (n_frames) *n_frames = s.n_frames;
    if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, s.ev[0], s.ev[5]);
    return 0;
}

extern "C" const uint32_t* b200flac_encoder_slot_frame_bytes(b200flac_encoder* enc, int slot)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) return nullptr;
    return enc->slots[slot].h_frame_bytes;
}

extern "C" const uint32_t* b200flac_encoder_slot_frame_pcm(b200flac_encoder* enc, int slot)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) return nullptr;
    return enc->slots[slot].frame_pcm.data();
}

extern "C" int b200flac_encoder_encode_device(b200flac_encoder* enc, int slot, const void* d_pcm,
                                              const b200flac_segment* segments, uint32_t n_segments,
                                              void* d_out, uint64_t out_capacity, uint64_t* out_bytes,
                                              uint32_t* n_frames, float* elapsed_ms)
{
    if (b200flac_encoder_submit_device(enc, slot, d_pcm, segments, n_segments, d_out, out_capacity)) return 1;
    return b200flac_encoder_collect_device(enc, slot, out_bytes, n_frames, elapsed_ms);
}

extern "C" int b200flac_encoder_last_kernel_ms(b200flac_encoder* enc, int slot, float* ms, int capacity)
{
    if (!enc || slot < 0 || slot >= enc->n_slots || !ms) return 0;
    Slot& s = enc->slots[slot];
    if (!s.timed) return 0;
    cudaSetDevice(enc->device);
    if (cudaEventSynchronize(s.ev[5]) != cudaSuccess) return 0;
    int n = 0;
    for (int i = 0; i < 5 && i < capacity; i++, n++) {
        ms[i] = 0.f;
        cudaEventElapsedTime(&ms[i], s.ev[i], s.ev[i + 1]);
    }
    return n;
}

extern "C" int b200flac_encoder_set_chunking(b200flac_encoder* enc, uint32_t chunk_frames, uint32_t lookahead)
{
    if (!enc) { set_err("encoder is NULL"); return 1; }
    for (int i = 0; i < enc->n_slots; i++) if (enc->slots[i].busy) { set_err("a slot is busy"); return 1; }
    enc->chunk_frames = chunk_frames;
    enc->lookahead = lookahead ? lookahead : 1;
    if (getenv("B200FLAC_NH")) enc->model_streams = (u32)atoi(getenv("B200FLAC_NH"));            // tuning knobs
    if (getenv("B200FLAC_LPC_PRIO")) enc->model_priority = (u32)atoi(getenv("B200FLAC_LPC_PRIO"));
    if (getenv("B200FLAC_LPC_GRID")) enc->lpc_grid_cap = (u32)atoi(getenv("B200FLAC_LPC_GRID"));
    return 0;
}

extern "C" uint64_t b200flac_encoder_launch_count(const b200flac_encoder* enc) { return enc ? enc->launches : 0; }

extern "C" int b200flac_encoder_get_plans(b200flac_encoder* enc, int slot, b200flac_plan* plans, uint8_t* rice,
                                          uint32_t* rice_stride, uint8_t* assignments, uint32_t* candidates_per_frame)
{
    if (!enc || slot < 0 || slot >= enc->n_slots) { set_err("bad encoder or slot"); return 1; }
    Slot& s = enc->slots[slot];
    CU_CHECK(cudaSetDevice(enc->device), 1);
    CU_CHECK(cudaStreamSynchronize(s.stream), 1);
    const size_t U = (size_t)s.n_frames * enc->P.K;
    if (plans) CU_CHECK(cudaMemcpy(plans, s.d_plans, U * sizeof(b200flac_plan), cudaMemcpyDeviceToHost), 1);
    if (rice) CU_CHECK(cudaMemcpy(rice, s.d_rice, U * enc->P.rice_stride, cudaMemcpyDeviceToHost), 1);
    if (rice_stride) *rice_stride = enc->P.rice_stride;
    if (candidates_per_frame) *candidates_per_frame = enc->P.K;
    if (assignments) {
        std::vector<bf_frame_choice> ch(s.n_frames);
        CU_CHECK(cudaMemcpy(ch.data(), s.d_choice, s.n_frames * sizeof(bf_frame_choice), cudaMemcpyDeviceToHost), 1);
        for (u32 i = 0; i < s.n_frames; i++) assignments[i] = ch[i].assignment;
    }
    return 0;
}
