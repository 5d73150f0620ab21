// b200flac_batch.cu -- many tracks -> many files in ONE call (include/b200flac.h, b200flac_encode_files).
//
// BASELINE.json config 5 is "a batch of 10,000 three-minute tracks encoded concurrently".  The reference has no such
// entry: FlacAudio.from_pcm (audiotools/flac.py:1235-1330) calls encoders.encode_flac once per file, and so can a
// user of this engine -- one host thread per track through b200flac_encode_file.  That path is bound by the
// STREAMINFO MD5 (flac.c:187-188, 277): one core hashes ~640 MB/s, a three-minute track takes it 50 ms, sixteen
// cores sign 320 tracks per second while the GPU is a few per cent busy.  Here the whole batch is one job:
//   * tracks are packed into batches of ~512 MB of PCM, copied to a ring of device regions straight from the
//     caller's memory, and encoded as many-segment batches of the frame layer (one segment per track);
//   * the MD5 of every track is computed ON THE DEVICE from the PCM that is there anyway: MD5 is serial inside a
//     stream but the streams are independent -- one thread per track, a warp hashes 32 tracks in lockstep.  A track
//     takes a thread ~0.3 s, so the regions stay alive in a ring until their kernel is done -- and the END of the
//     list, which the device could not hash in time, is hashed by the pool's spare host threads, sixteen tracks at a
//     time in the lanes of one vector (md5_lanes.cpp);
//   * frames come back through three pinned buffers and a pool of host threads writes the files (head, frames);
//     the 16 bytes of MD5 are patched into each STREAMINFO at the end.
// Every file is byte for byte what b200flac_encode_file writes for the same track (tests/test_gpu_parity.py).
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <deque>
#include <vector>

#include "../../include/b200flac.h"

extern "C" void b200flac_internal_set_error(const char* msg);                       // b200flac_encoder.cu
extern "C" void b200flac_internal_md5(const uint8_t* p, size_t n, uint8_t out[16]);  // b200flac_stream.cu (one string, scalar)
extern "C" int b200flac_internal_md5_many(const uint8_t* const* ptr, const uint64_t* nbytes, uint32_t n, uint8_t* digests,
                                          uint64_t piece_bytes, int (*between)(void*), void* arg);   // md5_lanes.cpp
void b200flac_internal_stream_head(const b200flac_params* params, uint32_t padding_size, const char* version,
                                   uint32_t min_frame, uint32_t max_frame, uint64_t total_samples,
                                   const uint8_t md5[16], std::vector<uint8_t>& head);  // b200flac_stream.cu

typedef unsigned int u32;
typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------------------------
// MD5 (RFC 1321) of many byte strings at once: thread t hashes bytes [off[t], off[t] + len[t]) of `base`
// (off 16-byte aligned).  Same round formulation as the host's md5_block (b200flac_stream.cu).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 md5_rol(u32 x, int s) { return __funnelshift_l(x, x, s); }
__device__ __forceinline__ u32 md5_pre(u32 w, u32 xi, u32 k)
{
    u32 r;
    asm("{\n\t.reg .u32 t;\n\tadd.u32 t, %1, %2;\n\tadd.u32 %0, t, %3;\n\t}" : "=r"(r) : "r"(xi), "r"(k), "r"(w));
    return r;
}

__device__ __forceinline__ void md5_rounds(u32& a, u32& b, u32& c, u32& d, const u32 (&x)[16])
{
    const u32 a0 = a, b0 = b, c0 = c, d0 = d;
#define MF1(x_, y, z) (z ^ (x_ & (y ^ z)))
#define MF2(x_, y, z) ((x_ & z) + (y & ~z))
#define MF3(x_, y, z) (x_ ^ (y ^ z))
#define MF4(x_, y, z) (y ^ (x_ | ~z))
// w + x[i] + K does not depend on the previous step: it is formed by an asm the compiler cannot reassociate into the
// chain, which is then LOP3 -> IADD3 -> LEA.HI (rotate and add in one) per step; left to itself the compiler
// put x[i] and K behind the round function in two dependent adds (4 instructions deep, 76 MB/s per thread)
#define MSTEP(f, w, x_, y, z, xi, k, s) (w = md5_pre(w, xi, k) + f(x_, y, z), w = md5_rol(w, s) + x_)
    MSTEP(MF1, a, b, c, d, x[0], 0xd76aa478u, 7);   MSTEP(MF1, d, a, b, c, x[1], 0xe8c7b756u, 12);
    MSTEP(MF1, c, d, a, b, x[2], 0x242070dbu, 17);  MSTEP(MF1, b, c, d, a, x[3], 0xc1bdceeeu, 22);
    MSTEP(MF1, a, b, c, d, x[4], 0xf57c0fafu, 7);   MSTEP(MF1, d, a, b, c, x[5], 0x4787c62au, 12);
    MSTEP(MF1, c, d, a, b, x[6], 0xa8304613u, 17);  MSTEP(MF1, b, c, d, a, x[7], 0xfd469501u, 22);
    MSTEP(MF1, a, b, c, d, x[8], 0x698098d8u, 7);   MSTEP(MF1, d, a, b, c, x[9], 0x8b44f7afu, 12);
    MSTEP(MF1, c, d, a, b, x[10], 0xffff5bb1u, 17); MSTEP(MF1, b, c, d, a, x[11], 0x895cd7beu, 22);
    MSTEP(MF1, a, b, c, d, x[12], 0x6b901122u, 7);  MSTEP(MF1, d, a, b, c, x[13], 0xfd987193u, 12);
    MSTEP(MF1, c, d, a, b, x[14], 0xa679438eu, 17); MSTEP(MF1, b, c, d, a, x[15], 0x49b40821u, 22);
    MSTEP(MF2, a, b, c, d, x[1], 0xf61e2562u, 5);   MSTEP(MF2, d, a, b, c, x[6], 0xc040b340u, 9);
    MSTEP(MF2, c, d, a, b, x[11], 0x265e5a51u, 14); MSTEP(MF2, b, c, d, a, x[0], 0xe9b6c7aau, 20);
    MSTEP(MF2, a, b, c, d, x[5], 0xd62f105du, 5);   MSTEP(MF2, d, a, b, c, x[10], 0x02441453u, 9);
    MSTEP(MF2, c, d, a, b, x[15], 0xd8a1e681u, 14); MSTEP(MF2, b, c, d, a, x[4], 0xe7d3fbc8u, 20);
    MSTEP(MF2, a, b, c, d, x[9], 0x21e1cde6u, 5);   MSTEP(MF2, d, a, b, c, x[14], 0xc33707d6u, 9);
    MSTEP(MF2, c, d, a, b, x[3], 0xf4d50d87u, 14);  MSTEP(MF2, b, c, d, a, x[8], 0x455a14edu, 20);
    MSTEP(MF2, a, b, c, d, x[13], 0xa9e3e905u, 5);  MSTEP(MF2, d, a, b, c, x[2], 0xfcefa3f8u, 9);
    MSTEP(MF2, c, d, a, b, x[7], 0x676f02d9u, 14);  MSTEP(MF2, b, c, d, a, x[12], 0x8d2a4c8au, 20);
    MSTEP(MF3, a, b, c, d, x[5], 0xfffa3942u, 4);   MSTEP(MF3, d, a, b, c, x[8], 0x8771f681u, 11);
    MSTEP(MF3, c, d, a, b, x[11], 0x6d9d6122u, 16); MSTEP(MF3, b, c, d, a, x[14], 0xfde5380cu, 23);
    MSTEP(MF3, a, b, c, d, x[1], 0xa4beea44u, 4);   MSTEP(MF3, d, a, b, c, x[4], 0x4bdecfa9u, 11);
    MSTEP(MF3, c, d, a, b, x[7], 0xf6bb4b60u, 16);  MSTEP(MF3, b, c, d, a, x[10], 0xbebfbc70u, 23);
    MSTEP(MF3, a, b, c, d, x[13], 0x289b7ec6u, 4);  MSTEP(MF3, d, a, b, c, x[0], 0xeaa127fau, 11);
    MSTEP(MF3, c, d, a, b, x[3], 0xd4ef3085u, 16);  MSTEP(MF3, b, c, d, a, x[6], 0x04881d05u, 23);
    MSTEP(MF3, a, b, c, d, x[9], 0xd9d4d039u, 4);   MSTEP(MF3, d, a, b, c, x[12], 0xe6db99e5u, 11);
    MSTEP(MF3, c, d, a, b, x[15], 0x1fa27cf8u, 16); MSTEP(MF3, b, c, d, a, x[2], 0xc4ac5665u, 23);
    MSTEP(MF4, a, b, c, d, x[0], 0xf4292244u, 6);   MSTEP(MF4, d, a, b, c, x[7], 0x432aff97u, 10);
    MSTEP(MF4, c, d, a, b, x[14], 0xab9423a7u, 15); MSTEP(MF4, b, c, d, a, x[5], 0xfc93a039u, 21);
    MSTEP(MF4, a, b, c, d, x[12], 0x655b59c3u, 6);  MSTEP(MF4, d, a, b, c, x[3], 0x8f0ccc92u, 10);
    MSTEP(MF4, c, d, a, b, x[10], 0xffeff47du, 15); MSTEP(MF4, b, c, d, a, x[1], 0x85845dd1u, 21);
    MSTEP(MF4, a, b, c, d, x[8], 0x6fa87e4fu, 6);   MSTEP(MF4, d, a, b, c, x[15], 0xfe2ce6e0u, 10);
    MSTEP(MF4, c, d, a, b, x[6], 0xa3014314u, 15);  MSTEP(MF4, b, c, d, a, x[13], 0x4e0811a1u, 21);
    MSTEP(MF4, a, b, c, d, x[4], 0xf7537e82u, 6);   MSTEP(MF4, d, a, b, c, x[11], 0xbd3af235u, 10);
    MSTEP(MF4, c, d, a, b, x[2], 0x2ad7d2bbu, 15);  MSTEP(MF4, b, c, d, a, x[9], 0xeb86d391u, 21);
#undef MSTEP
#undef MF1
#undef MF2
#undef MF3
#undef MF4
    a += a0; b += b0; c += c0; d += d0;
}

__global__ void __launch_bounds__(32)
k_md5_tracks(const uint8_t* __restrict__ base, const u64* __restrict__ off, const u64* __restrict__ len, u32 n_tracks,
             uint8_t* __restrict__ digests)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const uint4* p = (const uint4*)(base + off[t]);
    const u64 nbytes = len[t];
    const u64 nblocks = nbytes >> 6;
    u32 a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    // the next block's four loads are in flight while this one is hashed
    uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0;
    if (nblocks) { n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3); }
    // ... and the lines eight blocks further on are asked into L2: every lane walks its own track, so each load is a
    // DRAM access of its own, and one block of hashing (~1,000 cycles) does not cover it (ncu: the warp waited on the
    // long scoreboard half of its time)
    const u64 AHEAD = 8;
#pragma unroll 1
    for (u64 i = 0; i < nblocks; i++) {
        const u32 x[16] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, n3.x, n3.y, n3.z, n3.w};
        if (i + 1 + AHEAD < nblocks) asm volatile("prefetch.global.L2 [%0];" :: "l"(p + 4 * (i + 1 + AHEAD)));
        if (i + 1 < nblocks) {
            const uint4* q = p + 4 * (i + 1);
            n0 = __ldg(q); n1 = __ldg(q + 1); n2 = __ldg(q + 2); n3 = __ldg(q + 3);
        }
        md5_rounds(a, b, c, d, x);
    }
    // the tail: the remaining bytes, 0x80, zeros, the length in bits (one or two blocks)
    const uint8_t* tail = (const uint8_t*)(p + 4 * nblocks);
    const u32 rem = (u32)(nbytes & 63);
    uint8_t buf[128];
    for (u32 i = 0; i < 128; i++) buf[i] = 0;
    for (u32 i = 0; i < rem; i++) buf[i] = tail[i];
    buf[rem] = 0x80;
    const u32 tb = rem < 56 ? 1u : 2u;
    const u64 bits = nbytes << 3;
    for (u32 i = 0; i < 8; i++) buf[tb * 64 - 8 + i] = (uint8_t)(bits >> (8 * i));
    for (u32 k = 0; k < tb; k++) {
        u32 x[16];
        for (u32 i = 0; i < 16; i++)
            x[i] = (u32)buf[64 * k + 4 * i] | ((u32)buf[64 * k + 4 * i + 1] << 8) | ((u32)buf[64 * k + 4 * i + 2] << 16) |
                   ((u32)buf[64 * k + 4 * i + 3] << 24);
        md5_rounds(a, b, c, d, x);
    }
    u32* o = (u32*)(digests + 16 * (size_t)t);        // (may be mapped host memory)
    o[0] = a; o[1] = b; o[2] = c; o[3] = d;
    __threadfence_system();
}

// test hook: MD5 of one byte string on the device (tests compare it with hashlib)
extern "C" int b200flac_internal_device_md5(int device, const uint8_t* host_bytes, uint64_t n, uint8_t out[16])
{
    if (cudaSetDevice(device) != cudaSuccess) { b200flac_internal_set_error("cudaSetDevice failed"); return 1; }
    uint8_t* d = nullptr;
    u64* meta = nullptr;
    uint8_t* dg = nullptr;
    int rc = 1;
    if (cudaMalloc((void**)&d, n + 64) == cudaSuccess && cudaMalloc((void**)&meta, 16) == cudaSuccess &&
        cudaMalloc((void**)&dg, 16) == cudaSuccess) {
        const u64 h[2] = {0, n};
        cudaMemcpy(d, host_bytes, n, cudaMemcpyHostToDevice);
        cudaMemcpy(meta, h, 16, cudaMemcpyHostToDevice);
        k_md5_tracks<<<1, 32>>>(d, meta, meta + 1, 1, dg);
        if (cudaMemcpy(out, dg, 16, cudaMemcpyDeviceToHost) == cudaSuccess) rc = 0;
    }
    if (rc) b200flac_internal_set_error("device MD5 failed");
    cudaFree(d); cudaFree(meta); cudaFree(dg);
    return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// the batch job
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct Region {                 // one batch's PCM on the device + its hashes
    uint8_t* d_pcm;
    u64* d_meta;                // [2][cap_tracks]: byte offsets, byte lengths
    uint8_t* d_digest;          // [cap_tracks][16]
    u64* h_meta;                // pinned
    uint8_t* h_digest;          // pinned
    cudaStream_t st;            // its MD5 kernel's stream
    cudaEvent_t ev_h2d, ev_md5;
    int batch;                  // batch using it (-1: free)
    bool dev_md5;               // its hashes are being computed by the device and have not been read back
};

struct OutBuf {                 // frames of one batch on the host
    uint8_t* h;
    cudaEvent_t ev;
    int pending;                // tracks not yet written
};

const int NS = 3;               // batches in flight in the encoder (its slots; as many output buffers)

struct Ctx {                    // kept between calls (pinned memory is expensive to allocate)
    bool valid;
    b200flac_params params;
    int device;
    u64 batch_bytes, out_cap;
    u32 cap_tracks;
    b200flac_encoder* enc;
    std::vector<Region> ring;
    OutBuf out[NS];
    void* d_out[NS];
    cudaStream_t st_h2d, st_d2h;
};
Ctx g_ctx;
pthread_mutex_t g_ctx_mu = PTHREAD_MUTEX_INITIALIZER;

void ctx_destroy(Ctx& c)
{
    if (!c.valid) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (auto& r : c.ring) {
        cudaFree(r.d_pcm); cudaFree(r.d_meta); cudaFree(r.d_digest);
        cudaFreeHost(r.h_meta); cudaFreeHost(r.h_digest);
        cudaStreamDestroy(r.st); cudaEventDestroy(r.ev_h2d); cudaEventDestroy(r.ev_md5);
    }
    c.ring.clear();
    for (int i = 0; i < NS; i++) {
        if (c.out[i].h) cudaFreeHost(c.out[i].h);
        if (c.out[i].ev) cudaEventDestroy(c.out[i].ev);
        if (c.d_out[i]) cudaFree(c.d_out[i]);
        c.out[i].h = nullptr; c.out[i].ev = nullptr; c.d_out[i] = nullptr;
    }
    if (c.st_h2d) cudaStreamDestroy(c.st_h2d);
    if (c.st_d2h) cudaStreamDestroy(c.st_d2h);
    if (c.enc) b200flac_encoder_destroy(c.enc);
    c.enc = nullptr; c.st_h2d = c.st_d2h = nullptr;
    c.valid = false;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char m_[256]; \
    snprintf(m_, sizeof(m_), "%s: %s", #call, cudaGetErrorString(e_)); b200flac_internal_set_error(m_); return 1; } } while (0)

int ctx_prepare(Ctx& c, const b200flac_params* p, int device, u64 batch_bytes, u32 cap_tracks, size_t ring_regions)
{
    if (c.valid && (memcmp(&c.params, p, sizeof(*p)) != 0 || c.device != device || c.batch_bytes != batch_bytes ||
                    c.cap_tracks < cap_tracks))
        ctx_destroy(c);
    CK(cudaSetDevice(device));
    if (!c.valid) {
        memset(&c.params, 0, sizeof(c.params));
        c.params = *p; c.device = device; c.batch_bytes = batch_bytes; c.cap_tracks = cap_tracks;
        c.enc = nullptr; c.st_h2d = c.st_d2h = nullptr;
        for (int i = 0; i < NS; i++) { c.out[i].h = nullptr; c.out[i].ev = nullptr; c.out[i].pending = 0; c.d_out[i] = nullptr; }
        c.valid = true;
        const u64 frame_bytes = (u64)p->channels * (p->bits_per_sample / 8);
        const u64 batch_frames = batch_bytes / frame_bytes;
        c.enc = b200flac_encoder_create(p, device, batch_frames + (u64)cap_tracks * p->block_size, NS);
        if (!c.enc) { ctx_destroy(c); return 1; }
        // every track ends with a short block at worst: cap_tracks more frames than the PCM alone needs
        c.out_cap = b200flac_encoder_output_bound(c.enc, batch_frames, cap_tracks) + 64;
        CK(cudaStreamCreateWithFlags(&c.st_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.st_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < NS; i++) {
            CK(cudaMalloc(&c.d_out[i], c.out_cap));
            CK(cudaMallocHost((void**)&c.out[i].h, c.out_cap));
            CK(cudaEventCreateWithFlags(&c.out[i].ev, cudaEventDisableTiming));
        }
    }
    while (c.ring.size() < ring_regions) {
        Region r;
        memset(&r, 0, sizeof(r));
        r.batch = -1;
        CK(cudaMalloc((void**)&r.d_pcm, batch_bytes + 256));
        CK(cudaMalloc((void**)&r.d_meta, 2 * sizeof(u64) * c.cap_tracks));
        r.d_digest = nullptr;
        CK(cudaMallocHost((void**)&r.h_meta, 2 * sizeof(u64) * c.cap_tracks));
        CK(cudaMallocHost((void**)&r.h_digest, 16 * (size_t)c.cap_tracks));
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CK(cudaStreamCreateWithPriority(&r.st, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&r.ev_h2d, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&r.ev_md5, cudaEventDisableTiming));
        c.ring.push_back(r);
    }
    return 0;
}

struct TrackOut {               // what the writers need for one track
    u32 track;
    const uint8_t* frames;      // in the batch's pinned buffer
    u64 bytes;
    u32 min_frame, max_frame;
    int outbuf;
};

struct Job {
    const char* const* filenames;
    const b200flac_params* params;
    uint32_t padding_size;
    const char* version;
    const uint64_t* n_pcm_frames;
    // writer pool
    pthread_mutex_t mu;
    pthread_cond_t cv_task, cv_done;
    std::deque<TrackOut> tasks;
    bool quit, failed;
    Ctx* ctx;
    // The hashing is shared.  The device takes whole batches from the front of the list as their PCM arrives (a track
    // costs one of its threads ~0.3 s whatever else happens, a thousand tracks at once cost the same 0.3 s); pool
    // threads with no file to write take up to sixteen tracks at a time from the END of the list and hash them side by
    // side out of the caller's memory (md5_lanes.cpp: several GB/s per thread, but there are only host_threads of them
    // and they compete with the copies for the host's memory).  The device's share ends at the first batch that
    // reaches into the host's zone (below) -- so the last batches of a long job never leave the whole job waiting
    // for one slow device thread, and a job of a few long tracks is hashed by the host alone.
    bool host_md5;
    const uint8_t* const* pcm;
    u64 frame_bytes;
    uint8_t* digests;
    const u32* batch_of;        // per track
    char* host_touched;         // per batch: the host has claimed one of its tracks
    long long host_cursor;      // the next track the host claims (counts down)
    long long dev_end_track;    // tracks [0, dev_end_track) are the device's
    bool dev_stopped;
    // the host's zone: tracks with at most zone_bytes of PCM from themselves to the end of the list -- what arrives on
    // the device during the last (one device hash) of the job and could not be hashed there in time.  The pool takes
    // nothing in front of it while the device is still claiming: hashing more than that on the host only takes
    // memory bandwidth from the copies (measured: 696 of 1,000 tracks hashed by the pool, copies 15 % slower).
    const u64* suffix_bytes;    // per track: PCM bytes of tracks t .. n - 1
    u64 zone_bytes;
    int host_inflight;
    u64 host_bytes; double host_busy_s;   // hashed by the pool so far; seconds its threads spent on it
    u32 host_tracks;
    double t_start, write_busy_s;
    u64 written_bytes, written_pcm_bytes;
    int n_threads;
};

bool write_track(Job* j, const TrackOut& t, std::vector<uint8_t>& head)
{
    static const uint8_t zero[16] = {0};
    FILE* f = fopen(j->filenames[t.track], "wb");
    if (!f) return false;
    bool ok = true;
    b200flac_internal_stream_head(j->params, j->padding_size, j->version, t.min_frame, t.max_frame,
                                  j->n_pcm_frames[t.track], zero, head);
    if (fwrite(head.data(), 1, head.size(), f) != head.size()) ok = false;
    if (ok && t.bytes && fwrite(t.frames, 1, (size_t)t.bytes, f) != (size_t)t.bytes) ok = false;
    if (fclose(f) != 0) ok = false;
    return ok;
}

// (mutex held on entry and exit) write every queued file
void drain_writes(Job* j, std::vector<uint8_t>& head)
{
    while (!j->tasks.empty()) {
        const TrackOut t = j->tasks.front();
        j->tasks.pop_front();
        pthread_mutex_unlock(&j->mu);
        const double t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        const bool ok = write_track(j, t, head);
        const double t1 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        pthread_mutex_lock(&j->mu);
        j->write_busy_s += t1 - t0;
        j->written_bytes += t.bytes; j->written_pcm_bytes += j->n_pcm_frames[t.track] * j->frame_bytes;
        if (!ok) j->failed = true;
        if (--j->ctx->out[t.outbuf].pending == 0) pthread_cond_broadcast(&j->cv_done);
    }
}

struct Between { Job* j; std::vector<uint8_t>* head; double spent; };

// between two pieces of hashing: the files that have become ready come first
int between_pieces(void* arg)
{
    Between* b = (Between*)arg;
    pthread_mutex_lock(&b->j->mu);
    if (!b->j->tasks.empty()) {
        const double t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        drain_writes(b->j, *b->head);
        b->spent += std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0;
    }
    const bool stop = b->j->quit;
    pthread_mutex_unlock(&b->j->mu);
    return stop ? 1 : 0;
}

void* writer_main(void* arg)
{
    Job* j = (Job*)arg;
    std::vector<uint8_t> head;
    Between btw = {j, &head, 0.0};
    const u64 PIECE = 256u << 10;           // per lane: a hashing thread looks for files to write every ~4 MB
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    pthread_mutex_lock(&j->mu);
    for (;;) {
        if (!j->tasks.empty()) { drain_writes(j, head); continue; }
        if (j->host_md5 && !j->quit && j->host_cursor >= j->dev_end_track &&
            (j->dev_stopped || j->suffix_bytes[j->host_cursor] <= j->zone_bytes)) {
            // up to sixteen tracks from the end of the list, hashed side by side in the lanes of one vector
            // (md5_lanes.cpp: the chain of one track is as slow as ever, sixteen cost the same)
            const uint8_t* ptr[16];
            uint64_t len[16];
            u32 first = 0, n = 0;
            u64 nb = 0;
            // (a fair share of what is there: sixteen lanes cost the time of one, but one lane alone is three times
            // slower than the scalar code -- a job of a few long tracks gives every thread one track)
            const long long avail = j->host_cursor - j->dev_end_track + 1;
            const u32 share = (u32)std::min<long long>(16, std::max<long long>(1, avail / j->n_threads));
            while (n < share && j->host_cursor >= j->dev_end_track &&
                   (j->dev_stopped || j->suffix_bytes[j->host_cursor] <= j->zone_bytes)) {
                const u32 t = (u32)j->host_cursor--;
                j->host_touched[j->batch_of[t]] = 1;
                first = t; n++;
                nb += j->n_pcm_frames[t] * j->frame_bytes;
            }
            for (u32 i = 0; i < n; i++) { ptr[i] = j->pcm[first + i]; len[i] = j->n_pcm_frames[first + i] * j->frame_bytes; }
            j->host_inflight++;
            pthread_mutex_unlock(&j->mu);
            const double t0 = now();
            btw.spent = 0;
            if (n >= 4) b200flac_internal_md5_many(ptr, len, n, j->digests + 16 * (size_t)first, PIECE, between_pieces, &btw);
            else
                for (u32 i = 0; i < n; i++) {
                    b200flac_internal_md5(ptr[i], (size_t)len[i], j->digests + 16 * (size_t)(first + i));
                    if (between_pieces(&btw)) break;
                }
            const double busy = now() - t0 - btw.spent;     // (hashing alone: the files written in between are not its time)
            pthread_mutex_lock(&j->mu);
            j->host_inflight--;
            j->host_bytes += nb; j->host_busy_s += busy; j->host_tracks += n;
            pthread_cond_broadcast(&j->cv_done);
            continue;
        }
        if (j->quit) break;
        pthread_cond_wait(&j->cv_task, &j->mu);
    }
    pthread_mutex_unlock(&j->mu);
    return nullptr;
}

} // namespace

// The host's zone of a many-files job (see Job): how many bytes at the END of the track list the pool hashes.
// With the copies arriving at `pace` bytes/s, a batch is hashed on the device dev_hash_s after it arrives, so the
// device is done in time with everything but the last pace * dev_hash_s bytes; the pool takes those if it can hash
// them by then -- at hash_rate per thread, with the threads that writing the files (which comes first: all_bytes *
// ratio bytes at write_rate per thread) leaves over -- and otherwise as much as makes the two finish together.
// Defaults before a job has measurements of its own: pace 36 GB/s (a B200's inbound copies next to outbound ones
// and a busy pool), 2 GB/s of file per thread, compressed ratio 0.7, 2.5 GB/s of hashing per thread.
extern "C" uint64_t b200flac_internal_host_zone(uint64_t all_bytes, double pace, double dev_hash_s, int host_threads,
                                                double write_rate, double ratio, double hash_rate)
{
    const double t_run = (double)all_bytes / pace;
    const double write_s = (double)all_bytes * ratio / write_rate;                 // thread-seconds
    const double spare = (double)host_threads - write_s / (t_run + dev_hash_s);    // threads left for hashing
    if (spare <= 0) return 0;
    const double hr = 0.8 * hash_rate * spare;
    return (uint64_t)std::min(pace * dev_hash_s, (t_run + dev_hash_s) / (1.0 / hr + 1.0 / pace));
}

// who hashed the tracks of the last job, and at what rates (bench.py reports it next to the job's time)
static double g_last_stats[6];
extern "C" void b200flac_internal_batch_stats(double out[6])
{
    pthread_mutex_lock(&g_ctx_mu);
    for (int i = 0; i < 6; i++) out[i] = g_last_stats[i];
    pthread_mutex_unlock(&g_ctx_mu);
}

extern "C" void b200flac_internal_batch_clear(void)
{
    pthread_mutex_lock(&g_ctx_mu);
    ctx_destroy(g_ctx);
    pthread_mutex_unlock(&g_ctx_mu);
}

extern "C" int b200flac_encode_files(uint32_t n_tracks, const char* const* filenames, const b200flac_params* params,
                                     uint32_t padding_size, const char* version,
                                     const uint8_t* const* pcm, const uint64_t* n_pcm_frames,
                                     int device, int host_threads)
{
    if (!filenames || !params || !pcm || !n_pcm_frames) { b200flac_internal_set_error("NULL argument"); return 1; }
    if (b200flac_device_count() <= 0) {
        b200flac_internal_set_error("no CUDA device available: the B200 FLAC engine has no CPU fallback");
        return 1;
    }
    if (n_tracks == 0) return 0;
    if (device < 0) { const char* e = getenv("B200FLAC_DEVICE"); device = (e && *e) ? atoi(e) : 0; }
    if (host_threads <= 0) host_threads = 8;
    const u64 frame_bytes = (u64)params->channels * (params->bits_per_sample / 8);
    if (!frame_bytes || !params->block_size) { b200flac_internal_set_error("bad stream parameters"); return 1; }

    // ---- batches: consecutive tracks, up to batch_bytes of PCM each (a longer track is a batch of its own) ----
    u64 batch_bytes = 512ull << 20;
    { const char* e = getenv("B200FLAC_FILES_BATCH_MB"); if (e && atol(e) > 0) batch_bytes = (u64)atol(e) << 20; }
    // a track's region starts on a multiple of 16 PCM frames: a whole number of frames for the encoder's segment
    // offsets, 16-byte aligned for the hashing kernel's loads
    const u64 align = 16 * frame_bytes;
    auto padded = [&](u64 nbytes) -> u64 { return (nbytes + align - 1) / align * align; };
    u64 longest = 0;
    for (u32 t = 0; t < n_tracks; t++) longest = std::max<u64>(longest, padded(n_pcm_frames[t] * frame_bytes));
    if (longest > batch_bytes) batch_bytes = (longest + 0xFFFFF) & ~0xFFFFFull;
    struct Batch { u32 first, count; u64 bytes; u64 raw, longest; };   // bytes: padded; raw, longest: PCM bytes
    std::vector<Batch> batches;
    u32 cap_tracks = 1;
    {
        Batch b = {0, 0, 0, 0, 0};
        for (u32 t = 0; t < n_tracks; t++) {
            const u64 nb = n_pcm_frames[t] * frame_bytes;
            const u64 tb = padded(nb);
            if (b.count && (b.bytes + tb > batch_bytes || b.count >= 4096)) { batches.push_back(b); b = {t, 0, 0, 0, 0}; }
            b.count++; b.bytes += tb; b.raw += nb; b.longest = std::max(b.longest, nb);
        }
        batches.push_back(b);
        for (auto& x : batches) cap_tracks = std::max(cap_tracks, x.count);
    }
    const int NB = (int)batches.size();
    size_t ring_regions = 40;
    { const char* e = getenv("B200FLAC_FILES_RING"); if (e && atol(e) > 1) ring_regions = (size_t)atol(e); }
    // (more regions than batches in flight in the encoder, or the front could not move)
    ring_regions = std::min<size_t>(std::max<size_t>(ring_regions, (size_t)NS + 1), (size_t)NB);

    pthread_mutex_lock(&g_ctx_mu);
    Ctx& c = g_ctx;
    int rc = ctx_prepare(c, params, device, batch_bytes, cap_tracks, ring_regions);
    if (rc) { ctx_destroy(c); pthread_mutex_unlock(&g_ctx_mu); return 1; }
    for (auto& r : c.ring) { r.batch = -1; r.dev_md5 = false; }
    for (int i = 0; i < NS; i++) c.out[i].pending = 0;
    const int dbg = getenv("B200FLAC_FILES_DEBUG") ? atoi(getenv("B200FLAC_FILES_DEBUG")) : 0;   // timing experiments only -- 1: no hashing, 2: frames neither copied back nor written, 4: PCM copied once (region reuse without copies)
    std::vector<uint8_t> digests((size_t)n_tracks * 16, 0);
    std::vector<u32> batch_of((size_t)n_tracks);
    std::vector<char> host_touched(batches.size(), 0);
    u64 all_bytes = 0;
    for (size_t b = 0; b < batches.size(); b++) {
        for (u32 i = 0; i < batches[b].count; i++) batch_of[batches[b].first + i] = (u32)b;
        all_bytes += batches[b].raw;
    }

    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    Job job;
    job.t_start = t_start;
    job.filenames = filenames; job.params = params; job.padding_size = padding_size; job.version = version;
    job.n_pcm_frames = n_pcm_frames; job.quit = false; job.failed = false; job.ctx = &c;
    // B200FLAC_FILES_HOST_MD5=0: every hash on the device (A/B measurements)
    { const char* e = getenv("B200FLAC_FILES_HOST_MD5"); job.host_md5 = !(e && atoi(e) == 0) && !(dbg & 1); }
    job.pcm = pcm; job.frame_bytes = frame_bytes; job.digests = digests.data(); job.batch_of = batch_of.data();
    job.host_touched = host_touched.data(); job.host_cursor = (long long)n_tracks - 1; job.dev_end_track = 0;
    job.dev_stopped = false; job.host_inflight = 0;
    std::vector<u64> suffix_bytes((size_t)n_tracks + 1, 0);
    for (u32 t = n_tracks; t-- > 0;) suffix_bytes[t] = suffix_bytes[t + 1] + n_pcm_frames[t] * frame_bytes;
    job.suffix_bytes = suffix_bytes.data();
    job.host_bytes = 0; job.host_busy_s = 0; job.host_tracks = 0; job.write_busy_s = 0; job.n_threads = host_threads;
    job.written_bytes = 0; job.written_pcm_bytes = 0;
    // what a device thread hashes per second inside a job (measured on a B200; alone, with the chain at LOP3 -> IADD ->
    // LEA.HI per step and the lines ahead asked into L2, it makes 107-121 MB/s: tools/md5_rate.py)
    double dev_rate = 100e6;
    { const char* e = getenv("B200FLAC_FILES_DEV_MD5_MBS"); if (e && atof(e) > 0) dev_rate = atof(e) * 1e6; }
    // (until the copies have shown their pace: 36 GB/s, a B200's inbound copies next to outbound ones and a busy pool)
    u64 longest_track = 0;
    for (u32 t = 0; t < n_tracks; t++) longest_track = std::max<u64>(longest_track, n_pcm_frames[t] * frame_bytes);
    const double dev_hash_s = (double)longest_track / dev_rate;
    u64 arrived_bytes = 0;
    // how much of the end of the list is the pool's (mutex held): b200flac_internal_host_zone with the job's own
    // measurements once there are any
    const char* zone_forced = getenv("B200FLAC_FILES_HOST_ZONE_MB");     // tests: a fixed zone
    auto host_zone = [&]() -> u64 {
        if (zone_forced) return (u64)(atof(zone_forced) * 1048576.0);
        const double el = now() - job.t_start;
        const double pace = (el > 0.05 && arrived_bytes) ? (double)arrived_bytes / el : 36e9;
        const double write_rate = job.write_busy_s > 0.02 ? (double)job.written_bytes / job.write_busy_s : 2.0e9;
        const double ratio = job.written_pcm_bytes ? (double)job.written_bytes / (double)job.written_pcm_bytes : 0.7;
        const double hash_rate = job.host_busy_s > 0.02 ? (double)job.host_bytes / job.host_busy_s : 2.5e9;
        return b200flac_internal_host_zone(all_bytes, pace, dev_hash_s, host_threads, write_rate, ratio, hash_rate);
    };
    job.zone_bytes = host_zone();
    pthread_mutex_init(&job.mu, nullptr);
    pthread_cond_init(&job.cv_task, nullptr);
    pthread_cond_init(&job.cv_done, nullptr);
    std::vector<pthread_t> writers((size_t)host_threads);
    for (auto& w : writers) pthread_create(&w, nullptr, writer_main, &job);

    std::vector<b200flac_segment> segs;
    const u32 bs = params->block_size;
    // B200FLAC_FILES_TRACE=1: where the orchestrating thread waited, to stderr
    const bool trace = getenv("B200FLAC_FILES_TRACE") != nullptr;
    double w_ring = 0, w_h2d = 0, w_collect = 0, w_writers = 0, w_d2h = 0, w_tail_files = 0, w_tail_md5 = 0, w_submit = 0;
    std::vector<char> submitted((size_t)NB, 0);
    int next_h2d = 0;                      // next batch whose copy has not been issued
    bool fail = false;

    // Start hashing every batch whose copy has ARRIVED (asked, not waited for).  Two traps, both measured:
    //  * the digests are written by the kernel straight into page-locked host memory -- a device->host copy queued
    //    behind this (long) kernel would sit at the head of the copy engine's queue and hold every later copy of
    //    frames back until the kernel is done (63 ms per batch instead of 7);
    //  * the kernel is launched only once its input is there, never queued behind a cudaStreamWaitEvent on a copy
    //    that is still hundreds of milliseconds away: streams share a handful of hardware queues, and an entry
    //    waiting at the head of one holds back the encoder's kernels that happen to map to the same queue
    //    (the encoder ran at 20 ms per batch instead of 11).
    int next_md5 = 0;
    u32 dev_tracks = 0;
    // does the device hash batch b?  (see Job)
    auto claim_for_device = [&](int b) -> bool {
        const Batch& bb = batches[(size_t)b];
        bool dev = true;
        pthread_mutex_lock(&job.mu);
        if (job.host_md5) {
            arrived_bytes += bb.raw;
            job.zone_bytes = host_zone();
            // in front of the zone: the device's; the first batch that reaches into it ends the device's share
            dev = !job.dev_stopped && !job.host_touched[(size_t)b] && job.suffix_bytes[bb.first + bb.count - 1] > job.zone_bytes;
            if (!dev) job.dev_stopped = true;
            pthread_cond_broadcast(&job.cv_task);       // (the zone has moved, or the rest is the pool's)
        }
        if (dev) { job.dev_end_track = (long long)bb.first + bb.count; dev_tracks += bb.count; }
        pthread_mutex_unlock(&job.mu);
        return dev;
    };
    auto launch_ready_md5 = [&](int upto_must) -> int {
        while (next_md5 < next_h2d) {
            Region& r = c.ring[(size_t)next_md5 % c.ring.size()];
            if (next_md5 > upto_must && cudaEventQuery(r.ev_h2d) != cudaSuccess) break;
            if (next_md5 <= upto_must) CK(cudaEventSynchronize(r.ev_h2d));
            const Batch& bb = batches[(size_t)next_md5];
            r.dev_md5 = claim_for_device(next_md5);
            if (r.dev_md5) {
                if (!(dbg & 1))
                    k_md5_tracks<<<(bb.count + 31) / 32, 32, 0, r.st>>>(r.d_pcm, r.d_meta, r.d_meta + c.cap_tracks, bb.count, r.h_digest);
                CK(cudaGetLastError());
                CK(cudaEventRecord(r.ev_md5, r.st));
            }
            next_md5++;
        }
        return 0;
    };
    // copy batch b to its region of the ring.  must: the encoder needs
    // this batch now (otherwise the call gives up, returning 2, while the region's previous hashes are not done)
    auto issue_h2d = [&](int b, bool must) -> int {
        Region& r = c.ring[(size_t)b % c.ring.size()];
        if (r.batch >= 0) {
            // the region's previous batch: its encode has been collected; its hashes must be done, then kept
            if (r.batch >= next_md5) { if (!must) return 2; if (launch_ready_md5(r.batch)) return 1; }
            if (r.dev_md5) {
                if (!must && cudaEventQuery(r.ev_md5) != cudaSuccess) return 2;
                const double t0 = now();
                CK(cudaEventSynchronize(r.ev_md5));
                w_ring += now() - t0;
                const Batch& ob = batches[(size_t)r.batch];
                memcpy(&digests[(size_t)ob.first * 16], r.h_digest, (size_t)ob.count * 16);
                r.dev_md5 = false;
            }
        }
        r.batch = b;
        const Batch& bb = batches[(size_t)b];
        u64 off = 0;
        for (u32 i = 0; i < bb.count; i++) {
            const u32 t = bb.first + i;
            const u64 nb = n_pcm_frames[t] * frame_bytes;
            r.h_meta[i] = off;
            r.h_meta[c.cap_tracks + i] = nb;
            if (nb && !((dbg & 4) && b >= (int)c.ring.size())) CK(cudaMemcpyAsync(r.d_pcm + off, pcm[t], nb, cudaMemcpyHostToDevice, c.st_h2d));
            off += padded(nb);
        }
        CK(cudaMemcpyAsync(r.d_meta, r.h_meta, 2 * sizeof(u64) * c.cap_tracks, cudaMemcpyHostToDevice, c.st_h2d));
        CK(cudaEventRecord(r.ev_h2d, c.st_h2d));
        return 0;
    };
    // frames of batch b are in d_out[b % NS]: start bringing them to the host ...
    std::vector<TrackOut> outs_of[NS];
    auto start_d2h = [&](int b) -> int {
        const int ob = b % NS;
        uint64_t out_bytes = 0;
        uint32_t nfr = 0;
        double t0 = now();
        if (b200flac_encoder_collect_device(c.enc, ob, &out_bytes, &nfr, nullptr)) return 1;
        w_collect += now() - t0;
        if (out_bytes > c.out_cap) { b200flac_internal_set_error("encoded batch exceeds the output buffer"); return 1; }
        const uint32_t* fb = b200flac_encoder_slot_frame_bytes(c.enc, ob);
        // the host buffer still holds batch b - 2 until its files are written
        t0 = now();
        pthread_mutex_lock(&job.mu);
        while (c.out[ob].pending > 0) pthread_cond_wait(&job.cv_done, &job.mu);
        pthread_mutex_unlock(&job.mu);
        w_writers += now() - t0;
        if (!(dbg & 2)) CK(cudaMemcpyAsync(c.out[ob].h, c.d_out[ob], out_bytes, cudaMemcpyDeviceToHost, c.st_d2h));
        CK(cudaEventRecord(c.out[ob].ev, c.st_d2h));
        // per-track extents while the copy runs
        const Batch& bb = batches[(size_t)b];
        std::vector<TrackOut>& outs = outs_of[ob];
        outs.assign(bb.count, TrackOut());
        u64 pos = 0;
        u32 f = 0;
        for (u32 i = 0; i < bb.count; i++) {
            const u32 t = bb.first + i;
            const u32 nf = (u32)((n_pcm_frames[t] + bs - 1) / bs);
            TrackOut& o = outs[i];
            o.track = t; o.frames = c.out[ob].h + pos; o.bytes = 0; o.min_frame = 0xFFFFFF; o.max_frame = 0; o.outbuf = ob;
            for (u32 k = 0; k < nf; k++, f++) {
                const u32 sz = fb[f];
                o.bytes += sz;
                o.min_frame = std::min(o.min_frame, sz);
                o.max_frame = std::max(o.max_frame, sz);
            }
            pos += o.bytes;
        }
        if (f != nfr || pos != out_bytes) { b200flac_internal_set_error("internal: batch frame accounting"); return 1; }
        return 0;
    };
    // ... and, an iteration later, queue the file writes: the copy runs under the next batch's submit and the wait
    // for the encoder.  (Measured neutral on a 16-core box: with the pool writing files and hashing, the two copy
    // directions together make 59 GB/s and pace the job whichever of the two the thread waits for.)
    auto complete_d2h = [&](int b) -> int {
        const int ob = b % NS;
        std::vector<TrackOut>& outs = outs_of[ob];
        const double t0 = now();
        CK(cudaEventSynchronize(c.out[ob].ev));
        w_d2h += now() - t0;
        pthread_mutex_lock(&job.mu);
        c.out[ob].pending += (int)outs.size();
        if (dbg & 2) for (auto& o : outs) { o.bytes = 0; o.frames = nullptr; }
        for (auto& o : outs) job.tasks.push_back(o);
        pthread_cond_broadcast(&job.cv_task);
        pthread_mutex_unlock(&job.mu);
        return 0;
    };
    auto run = [&]() -> int {
        for (int b = 0; b < NB; b++) {
            // The copy-and-hash front runs as far ahead of the encoder as the ring allows: a track's hash takes a
            // thread some 0.3-0.5 s whatever else happens, so it has to start long before the track's frames are
            // due -- the last batch's above all, or the whole job waits for it at the end.  A region may be
            // refilled once the batch that used it has been collected (batches <= b - NS here) and hashed.
            const int R = (int)c.ring.size();
            while (next_h2d < NB && (next_h2d < R || next_h2d - R <= b - NS)) {
                const int rc_ = issue_h2d(next_h2d, next_h2d == b);
                if (rc_ == 1) return 1;
                if (rc_ == 2) break;
                next_h2d++;
            }
            Region& r = c.ring[(size_t)b % c.ring.size()];
            const Batch& bb = batches[(size_t)b];
            const int ob = b % NS;
            double t0 = now();
            CK(cudaEventSynchronize(r.ev_h2d));
            w_h2d += now() - t0;
            if (launch_ready_md5(b)) return 1;
            segs.clear();
            for (u32 i = 0; i < bb.count; i++) {
                b200flac_segment sg;
                sg.pcm_frame_offset = r.h_meta[i] / frame_bytes;
                sg.n_pcm_frames = n_pcm_frames[bb.first + i];
                sg.first_frame_number = 0; sg.reserved = 0;
                if (sg.n_pcm_frames) segs.push_back(sg);
            }
            // (the slot's output buffer on the device is free once the copy of batch b - NS has arrived)
            if (b >= NS && submitted[(size_t)(b - NS)] && complete_d2h(b - NS)) return 1;
            if (!segs.empty()) {
                t0 = now();
                if (b200flac_encoder_submit_device(c.enc, ob, r.d_pcm, segs.data(), (u32)segs.size(), c.d_out[ob], c.out_cap))
                    return 1;
                w_submit += now() - t0;
                submitted[(size_t)b] = 1;
            }
            if (b >= NS - 1 && submitted[(size_t)(b - (NS - 1))] && start_d2h(b - (NS - 1))) return 1;
            if (segs.empty()) {
                // a batch of empty tracks: nothing was submitted; give the writers their (frameless) files
                pthread_mutex_lock(&job.mu);
                c.out[ob].pending += (int)bb.count;
                for (u32 i = 0; i < bb.count; i++) {
                    TrackOut o; o.track = bb.first + i; o.frames = nullptr; o.bytes = 0; o.min_frame = 0xFFFFFF; o.max_frame = 0; o.outbuf = ob;
                    job.tasks.push_back(o);
                }
                pthread_cond_broadcast(&job.cv_task);
                pthread_mutex_unlock(&job.mu);
            }
        }
        // (the last batches; a batch of empty tracks was not submitted and its files are queued already)
        for (int b = NB; b < NB + NS; b++) {
            if (b >= NS && submitted[(size_t)(b - NS)] && complete_d2h(b - NS)) return 1;
            const int s_ = b - (NS - 1);
            if (s_ >= 0 && s_ < NB && submitted[(size_t)s_] && start_d2h(s_)) return 1;
        }
        return 0;
    };
    if (run()) fail = true;

    // ---- drain: files written, hashes read back, STREAMINFO patched ----
    const double t_run = now();
    pthread_mutex_lock(&job.mu);
    if (!fail) for (int i = 0; i < NS; i++) while (c.out[i].pending > 0) pthread_cond_wait(&job.cv_done, &job.mu);
    w_tail_files = now() - t_run;
    // (every batch has been offered to the device by now: what it did not take is the host's, down to the last track)
    if (!fail && job.host_md5)
        while (job.host_cursor >= job.dev_end_track || job.host_inflight > 0) pthread_cond_wait(&job.cv_done, &job.mu);
    job.quit = true;
    pthread_cond_broadcast(&job.cv_task);
    pthread_mutex_unlock(&job.mu);
    for (auto& w : writers) pthread_join(w, nullptr);
    if (job.failed) { b200flac_internal_set_error("cannot write an output file"); fail = true; }
    const double t_md5 = t_run + w_tail_files;
    if (!fail) {
        for (auto& r : c.ring) {
            if (r.batch < 0 || !r.dev_md5) { r.batch = -1; continue; }
            if (cudaEventSynchronize(r.ev_md5) != cudaSuccess) { b200flac_internal_set_error("device MD5 failed"); fail = true; break; }
            const Batch& ob = batches[(size_t)r.batch];
            memcpy(&digests[(size_t)ob.first * 16], r.h_digest, (size_t)ob.count * 16);
            r.batch = -1; r.dev_md5 = false;
        }
    }
    if (!fail) {
        for (u32 t = 0; t < n_tracks && !fail; t++) {
            FILE* f = fopen(filenames[t], "r+b");
            if (!f || fseek(f, 8 + 18, SEEK_SET) != 0 || fwrite(&digests[(size_t)t * 16], 1, 16, f) != 16) fail = true;
            if (f && fclose(f) != 0) fail = true;
            if (fail) b200flac_internal_set_error("cannot write an output file");
        }
    }
    w_tail_md5 = now() - t_md5;
    if (trace)
        fprintf(stderr, "b200flac_encode_files: %u tracks, %d batches, ring %zu: total %.3f s; waits: ring(md5) %.3f h2d %.3f "
                "submit %.3f collect %.3f writers %.3f d2h %.3f | tail: files %.3f md5+patch %.3f | hashed: device %u tracks, "
                "host %u (%.0f MB/s per thread, %.2f thread-s); files %.2f thread-s\n",
                n_tracks, NB, c.ring.size(), now() - t_start, w_ring, w_h2d, w_submit, w_collect, w_writers, w_d2h,
                w_tail_files, w_tail_md5, dev_tracks, job.host_tracks,
                job.host_busy_s > 0 ? job.host_bytes / job.host_busy_s / 1e6 : 0.0, job.host_busy_s, job.write_busy_s);
    g_last_stats[0] = dev_tracks; g_last_stats[1] = job.host_tracks;
    g_last_stats[2] = job.host_busy_s > 0 ? job.host_bytes / job.host_busy_s : 0.0;          // bytes per second per pool thread, hashing
    g_last_stats[3] = job.write_busy_s > 0 ? job.written_bytes / job.write_busy_s : 0.0;    // bytes per second per pool thread, writing files
    g_last_stats[4] = job.host_busy_s; g_last_stats[5] = job.write_busy_s;                  // thread-seconds
    pthread_mutex_destroy(&job.mu);
    pthread_cond_destroy(&job.cv_task);
    pthread_cond_destroy(&job.cv_done);
    if (fail) ctx_destroy(c);            // (buffers may still be in use by failed work: start clean next time)
    pthread_mutex_unlock(&g_ctx_mu);
    return fail ? 1 : 0;
}
