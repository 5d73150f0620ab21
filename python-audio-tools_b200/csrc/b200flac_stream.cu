// b200flac_stream.cu -- stream layer of the C ABI (include/b200flac.h).
//
// Host-side replacement of the C-signature entry point
//   encoders_encode_flac()                      src/encoders/flac.c:124-306
// stream head (flac.c:209-238), frame loop (flac.c:244-274) feeding batches of
// blocks to the device encoders, STREAMINFO MD5 on a host thread overlapped
// with the GPU work (flac.c:187-188, 277), STREAMINFO rewrite (flac.c:277-279).
//
// Batches go round-robin over (device, slot) lanes and are collected in the
// same order, so frames reach the file in stream order; no inter-GPU traffic.
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <deque>
#include <vector>

#include "../../include/b200flac.h"

extern "C" void b200flac_internal_set_error(const char* msg); // b200flac_encoder.cu
extern "C" int b200flac_internal_is_pinned(const void* p);    // b200flac_encoder.cu

// ---------------------------------------------------------------------------
// MD5 (RFC 1321).  The reference hashes the PCM as signed little-endian bytes
// through a pcmreader callback (flac.c:188, src/common/md5.c); here the same
// bytes are hashed from the pinned staging buffers by a worker thread.
// ---------------------------------------------------------------------------
struct Md5 {
    uint32_t a, b, c, d;
    uint64_t len;
    uint8_t buf[64];
    unsigned fill;
};

static inline uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

static void md5_block(Md5* m, const uint8_t* p)
{
    uint32_t x[16];
    memcpy(x, p, 64); // little-endian host
    uint32_t a = m->a, b = m->b, c = m->c, d = m->d;
    // x_ is the value produced by the previous step, so each round function is written with as little
    // work as possible after x_: F2's two terms have no bits in common and join the additions
    // ((x & z) + (y & ~z)), F3 folds y ^ z first, and w + data never waits for x_ at all
#define F1(x, y, z) (z ^ (x & (y ^ z)))
#define F2(x, y, z) ((x & z) + (y & ~z))
#define F3(x, y, z) (x ^ (y ^ z))
#define F4(x, y, z) (y ^ (x | ~z))
#define STEP(f, w, x_, y, z, data, s) (w = (w + (data)) + f(x_, y, z), w = rol(w, s) + x_)
    STEP(F1, a, b, c, d, x[0] + 0xd76aa478, 7);   STEP(F1, d, a, b, c, x[1] + 0xe8c7b756, 12);
    STEP(F1, c, d, a, b, x[2] + 0x242070db, 17);  STEP(F1, b, c, d, a, x[3] + 0xc1bdceee, 22);
    STEP(F1, a, b, c, d, x[4] + 0xf57c0faf, 7);   STEP(F1, d, a, b, c, x[5] + 0x4787c62a, 12);
    STEP(F1, c, d, a, b, x[6] + 0xa8304613, 17);  STEP(F1, b, c, d, a, x[7] + 0xfd469501, 22);
    STEP(F1, a, b, c, d, x[8] + 0x698098d8, 7);   STEP(F1, d, a, b, c, x[9] + 0x8b44f7af, 12);
    STEP(F1, c, d, a, b, x[10] + 0xffff5bb1, 17); STEP(F1, b, c, d, a, x[11] + 0x895cd7be, 22);
    STEP(F1, a, b, c, d, x[12] + 0x6b901122, 7);  STEP(F1, d, a, b, c, x[13] + 0xfd987193, 12);
    STEP(F1, c, d, a, b, x[14] + 0xa679438e, 17); STEP(F1, b, c, d, a, x[15] + 0x49b40821, 22);
    STEP(F2, a, b, c, d, x[1] + 0xf61e2562, 5);   STEP(F2, d, a, b, c, x[6] + 0xc040b340, 9);
    STEP(F2, c, d, a, b, x[11] + 0x265e5a51, 14); STEP(F2, b, c, d, a, x[0] + 0xe9b6c7aa, 20);
    STEP(F2, a, b, c, d, x[5] + 0xd62f105d, 5);   STEP(F2, d, a, b, c, x[10] + 0x02441453, 9);
    STEP(F2, c, d, a, b, x[15] + 0xd8a1e681, 14); STEP(F2, b, c, d, a, x[4] + 0xe7d3fbc8, 20);
    STEP(F2, a, b, c, d, x[9] + 0x21e1cde6, 5);   STEP(F2, d, a, b, c, x[14] + 0xc33707d6, 9);
    STEP(F2, c, d, a, b, x[3] + 0xf4d50d87, 14);  STEP(F2, b, c, d, a, x[8] + 0x455a14ed, 20);
    STEP(F2, a, b, c, d, x[13] + 0xa9e3e905, 5);  STEP(F2, d, a, b, c, x[2] + 0xfcefa3f8, 9);
    STEP(F2, c, d, a, b, x[7] + 0x676f02d9, 14);  STEP(F2, b, c, d, a, x[12] + 0x8d2a4c8a, 20);
    STEP(F3, a, b, c, d, x[5] + 0xfffa3942, 4);   STEP(F3, d, a, b, c, x[8] + 0x8771f681, 11);
    STEP(F3, c, d, a, b, x[11] + 0x6d9d6122, 16); STEP(F3, b, c, d, a, x[14] + 0xfde5380c, 23);
    STEP(F3, a, b, c, d, x[1] + 0xa4beea44, 4);   STEP(F3, d, a, b, c, x[4] + 0x4bdecfa9, 11);
    STEP(F3, c, d, a, b, x[7] + 0xf6bb4b60, 16);  STEP(F3, b, c, d, a, x[10] + 0xbebfbc70, 23);
    STEP(F3, a, b, c, d, x[13] + 0x289b7ec6, 4);  STEP(F3, d, a, b, c, x[0] + 0xeaa127fa, 11);
    STEP(F3, c, d, a, b, x[3] + 0xd4ef3085, 16);  STEP(F3, b, c, d, a, x[6] + 0x04881d05, 23);
    STEP(F3, a, b, c, d, x[9] + 0xd9d4d039, 4);   STEP(F3, d, a, b, c, x[12] + 0xe6db99e5, 11);
    STEP(F3, c, d, a, b, x[15] + 0x1fa27cf8, 16); STEP(F3, b, c, d, a, x[2] + 0xc4ac5665, 23);
    STEP(F4, a, b, c, d, x[0] + 0xf4292244, 6);   STEP(F4, d, a, b, c, x[7] + 0x432aff97, 10);
    STEP(F4, c, d, a, b, x[14] + 0xab9423a7, 15); STEP(F4, b, c, d, a, x[5] + 0xfc93a039, 21);
    STEP(F4, a, b, c, d, x[12] + 0x655b59c3, 6);  STEP(F4, d, a, b, c, x[3] + 0x8f0ccc92, 10);
    STEP(F4, c, d, a, b, x[10] + 0xffeff47d, 15); STEP(F4, b, c, d, a, x[1] + 0x85845dd1, 21);
    STEP(F4, a, b, c, d, x[8] + 0x6fa87e4f, 6);   STEP(F4, d, a, b, c, x[15] + 0xfe2ce6e0, 10);
    STEP(F4, c, d, a, b, x[6] + 0xa3014314, 15);  STEP(F4, b, c, d, a, x[13] + 0x4e0811a1, 21);
    STEP(F4, a, b, c, d, x[4] + 0xf7537e82, 6);   STEP(F4, d, a, b, c, x[11] + 0xbd3af235, 10);
    STEP(F4, c, d, a, b, x[2] + 0x2ad7d2bb, 15);  STEP(F4, b, c, d, a, x[9] + 0xeb86d391, 21);
#undef STEP
#undef F1
#undef F2
#undef F3
#undef F4
    m->a += a; m->b += b; m->c += c; m->d += d;
}

static void md5_init(Md5* m) { m->a = 0x67452301; m->b = 0xefcdab89; m->c = 0x98badcfe; m->d = 0x10325476; m->len = 0; m->fill = 0; }

static void md5_update(Md5* m, const uint8_t* p, size_t n)
{
    m->len += n;
    if (m->fill) {
        size_t take = 64 - m->fill;
        if (take > n) take = n;
        memcpy(m->buf + m->fill, p, take);
        m->fill += (unsigned)take; p += take; n -= take;
        if (m->fill == 64) { md5_block(m, m->buf); m->fill = 0; }
    }
    while (n >= 64) { md5_block(m, p); p += 64; n -= 64; }
    if (n) { memcpy(m->buf, p, n); m->fill = (unsigned)n; }
}

static void md5_final(Md5* m, uint8_t out[16])
{
    const uint64_t bits = m->len * 8;
    uint8_t pad[72];
    const size_t padlen = (m->fill < 56) ? (56 - m->fill) : (120 - m->fill);
    memset(pad, 0, sizeof(pad));
    pad[0] = 0x80;
    for (int i = 0; i < 8; i++) pad[padlen + i] = (uint8_t)(bits >> (8 * i));
    md5_update(m, pad, padlen + 8);
    const uint32_t h[4] = {m->a, m->b, m->c, m->d};
    for (int i = 0; i < 16; i++) out[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
}

// one-shot digest for the decode/verify layer (b200flac_decoder.cu)
extern "C" void b200flac_internal_md5(const uint8_t* p, size_t n, uint8_t out[16])
{
    Md5 m;
    md5_init(&m);
    md5_update(&m, p, n);
    md5_final(&m, out);
}

// ---------------------------------------------------------------------------
struct Lane {
    b200flac_encoder* enc;
    int slot;
    uint8_t* pcm;          // pinned staging of this lane
    const uint8_t* ext;    // this batch's PCM lies in the caller's own page-locked memory (no staging copy)
    uint64_t fill;         // PCM frames staged
    uint64_t md5_enq;      // PCM frames of this batch already handed to the MD5 thread
    std::vector<b200flac_segment> segs;
    uint64_t seg_start;    // first PCM frame of the open segment
    bool in_flight;
    uint64_t md5_ticket;   // MD5 job number that must complete before the staging is reused
};

struct Md5Job { const uint8_t* p; size_t n; };
static const size_t MD5_PIECE = 4u << 20; // bytes of a filling batch handed to the MD5 thread at a time

struct b200flac_stream {
    FILE* f;
    b200flac_params params;
    uint32_t padding_size;
    std::vector<b200flac_encoder*> encs;
    std::vector<int> enc_devices;
    int slots_per_dev;
    bool reusable;              // encoders may go back to the pool (no failure so far)
    std::vector<Lane> lanes;
    size_t cur;                 // lane being filled
    size_t oldest;              // next lane to collect
    size_t n_in_flight;
    uint64_t batch_frames;      // PCM frames per batch (multiple of block_size)
    uint32_t next_frame_number; // total_flac_frames, flac.c:205
    uint64_t total_samples;
    uint32_t min_frame, max_frame;
    uint64_t current_offset;
    std::vector<uint64_t> offsets;
    std::vector<uint32_t> lengths;
    std::vector<uint32_t> fbytes, fpcm;
    bool failed;
    // md5 worker
    pthread_t md5_thread;
    pthread_mutex_t mu;
    pthread_cond_t cv_job, cv_done;
    std::deque<Md5Job> jobs;
    uint64_t jobs_submitted, jobs_done;
    bool md5_quit;
    Md5 md5;
};

static void stream_err(const char* msg)
{
    if (msg != b200flac_last_error()) b200flac_internal_set_error(msg);
}

static void* md5_worker(void* arg)
{
    b200flac_stream* s = (b200flac_stream*)arg;
    pthread_mutex_lock(&s->mu);
    for (;;) {
        while (s->jobs.empty() && !s->md5_quit) pthread_cond_wait(&s->cv_job, &s->mu);
        if (s->jobs.empty() && s->md5_quit) break;
        Md5Job j = s->jobs.front();
        s->jobs.pop_front();
        pthread_mutex_unlock(&s->mu);
        md5_update(&s->md5, j.p, j.n);
        pthread_mutex_lock(&s->mu);
        s->jobs_done++;
        pthread_cond_broadcast(&s->cv_done);
    }
    pthread_mutex_unlock(&s->mu);
    return nullptr;
}

static void md5_wait(b200flac_stream* s, uint64_t ticket)
{
    pthread_mutex_lock(&s->mu);
    while (s->jobs_done < ticket) pthread_cond_wait(&s->cv_done, &s->mu);
    pthread_mutex_unlock(&s->mu);
}

static void put_be(uint8_t* p, uint64_t v, int bytes)
{
    for (int i = 0; i < bytes; i++) p[i] = (uint8_t)(v >> (8 * (bytes - 1 - i)));
}

// STREAMINFO body, flac.c:376-409
static void streaminfo_body(const b200flac_params& p, uint32_t min_frame, uint32_t max_frame, uint64_t total_samples,
                            const uint8_t md5[16], uint8_t out[34])
{
    auto clampu = [](uint64_t v, uint64_t hi) { return v > hi ? hi : v; };
    put_be(out + 0, clampu(p.block_size, 0xFFFF), 2);
    put_be(out + 2, clampu(p.block_size, 0xFFFF), 2);
    put_be(out + 4, clampu(min_frame, 0xFFFFFF), 3);
    put_be(out + 7, clampu(max_frame, 0xFFFFFF), 3);
    // 20 bits rate | 3 bits channels-1 | 5 bits bps-1 | 36 bits total samples
    const uint64_t v = (clampu(p.sample_rate, 0xFFFFF) << 44) | (clampu(p.channels - 1, 7) << 41) |
                       (clampu(p.bits_per_sample - 1, 31) << 36) | (total_samples & 0xFFFFFFFFFull);
    put_be(out + 10, v, 8);
    memcpy(out + 18, md5, 16);
}
static void build_streaminfo(const b200flac_stream* s, const uint8_t md5[16], uint8_t out[34])
{
    streaminfo_body(s->params, s->min_frame, s->max_frame, s->total_samples, md5, out);
}

// The bytes before the first frame, flac.c:209-238: "fLaC", STREAMINFO, VORBIS_COMMENT (vendor string only), PADDING.
// The STREAMINFO body is bytes 8..41 (the MD5 its last 16).  Shared with the many-files entry (b200flac_batch.cu).
void b200flac_internal_stream_head(const b200flac_params* params, uint32_t padding_size, const char* version,
                                   uint32_t min_frame, uint32_t max_frame, uint64_t total_samples,
                                   const uint8_t md5[16], std::vector<uint8_t>& head)
{
    char vendor[300];
    snprintf(vendor, sizeof(vendor), "Python Audio Tools %s", version ? version : "2.22alpha1");
    const uint32_t L = (uint32_t)strlen(vendor);
    head.clear();
    const uint8_t magic[4] = {0x66, 0x4C, 0x61, 0x43};
    head.insert(head.end(), magic, magic + 4);
    uint8_t bh[4];
    bh[0] = 0x00; put_be(bh + 1, 34, 3);                 // not last | STREAMINFO | 34
    head.insert(head.end(), bh, bh + 4);
    uint8_t si[34];
    streaminfo_body(*params, min_frame, max_frame, total_samples, md5, si);
    head.insert(head.end(), si, si + 34);
    bh[0] = 0x04; put_be(bh + 1, 4 + L + 4, 3);          // VORBIS_COMMENT
    head.insert(head.end(), bh, bh + 4);
    for (int i = 0; i < 4; i++) head.push_back((uint8_t)(L >> (8 * i))); // little-endian fields
    head.insert(head.end(), vendor, vendor + L);
    for (int i = 0; i < 4; i++) head.push_back(0);
    bh[0] = 0x81; put_be(bh + 1, padding_size, 3);       // last | PADDING
    head.insert(head.end(), bh, bh + 4);
    head.resize(head.size() + padding_size, 0);
}

// ---- encoder pool -------------------------------------------------------------------------------
// Creating an encoder costs far more than encoding a short file with it (pinned staging and device
// buffers for 2 x 2048 blocks: ~100 ms of cudaMallocHost/cudaMalloc against ~2 ms of kernels for ten
// minutes of audio), so idle encoders are kept for the next stream with the same options -- the
// 10,000-track batch of BASELINE.json config #5 pays the set-up once per worker, not once per file.
struct PoolEntry {
    b200flac_params params;
    int device;
    uint64_t batch_frames;
    int slots;
    b200flac_encoder* enc;
};
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;
static std::vector<PoolEntry> g_pool;
static size_t pool_capacity()
{
    // idle encoders kept (each: ~70 MB pinned, ~250 MB device at 16-bit stereo); B200FLAC_POOL=0 disables
    static long cap = -1;
    if (cap < 0) { const char* e = getenv("B200FLAC_POOL"); cap = e ? atol(e) : 16; if (cap < 0) cap = 0; }
    return (size_t)cap;
}

static b200flac_encoder* pool_acquire(const b200flac_params* p, int device, uint64_t batch_frames, int slots)
{
    b200flac_encoder* e = nullptr;
    pthread_mutex_lock(&g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); i++) {
        const PoolEntry& pe = g_pool[i];
        if (pe.device == device && pe.batch_frames == batch_frames && pe.slots == slots &&
            memcmp(&pe.params, p, sizeof(*p)) == 0) {
            e = pe.enc;
            g_pool.erase(g_pool.begin() + (long)i);
            break;
        }
    }
    pthread_mutex_unlock(&g_pool_mu);
    return e ? e : b200flac_encoder_create(p, device, batch_frames, slots);
}

static void pool_release(const b200flac_params* p, int device, uint64_t batch_frames, int slots, b200flac_encoder* e)
{
    b200flac_encoder* evict = nullptr;
    pthread_mutex_lock(&g_pool_mu);
    if (pool_capacity() == 0) { pthread_mutex_unlock(&g_pool_mu); b200flac_encoder_destroy(e); return; }
    if (g_pool.size() >= pool_capacity()) { evict = g_pool.front().enc; g_pool.erase(g_pool.begin()); }
    PoolEntry pe;
    memset(&pe, 0, sizeof(pe));
    pe.params = *p; pe.device = device; pe.batch_frames = batch_frames; pe.slots = slots; pe.enc = e;
    g_pool.push_back(pe);
    pthread_mutex_unlock(&g_pool_mu);
    if (evict) b200flac_encoder_destroy(evict);
}

extern "C" void b200flac_internal_decoder_clear(void); // b200flac_decoder.cu
extern "C" void b200flac_internal_batch_clear(void);   // b200flac_batch.cu

extern "C" void b200flac_pool_clear(void)
{
    b200flac_internal_decoder_clear();
    b200flac_internal_batch_clear();
    std::vector<PoolEntry> all;
    pthread_mutex_lock(&g_pool_mu);
    all.swap(g_pool);
    pthread_mutex_unlock(&g_pool_mu);
    for (auto& pe : all) b200flac_encoder_destroy(pe.enc);
}

static void destroy_stream(b200flac_stream* s)
{
    if (!s) return;
    pthread_mutex_lock(&s->mu);
    s->md5_quit = true;
    pthread_cond_broadcast(&s->cv_job);
    pthread_mutex_unlock(&s->mu);
    pthread_join(s->md5_thread, nullptr);
    for (size_t i = 0; i < s->encs.size(); i++) {
        // an encoder with nothing in flight goes back to the pool; after a failure or an abort with
        // batches in flight it is destroyed (destroy synchronises its streams)
        if (s->reusable && !s->failed && s->n_in_flight == 0) pool_release(&s->params, s->enc_devices[i], s->batch_frames, s->slots_per_dev, s->encs[i]);
        else b200flac_encoder_destroy(s->encs[i]);
    }
    if (s->f) fclose(s->f);
    pthread_mutex_destroy(&s->mu);
    pthread_cond_destroy(&s->cv_job);
    pthread_cond_destroy(&s->cv_done);
    delete s;
}

extern "C" b200flac_stream* b200flac_stream_open(const char* filename, const b200flac_params* params,
                                                 uint32_t padding_size, const char* version,
                                                 const int* devices, int n_devices)
{
    if (!filename || !params) { stream_err("filename/params is NULL"); return nullptr; }
    if (b200flac_device_count() <= 0) {
        stream_err("no CUDA device available: the B200 FLAC engine has no CPU fallback");
        return nullptr;
    }
    FILE* f = fopen(filename, "wb");
    if (!f) {
        char msg[400];
        snprintf(msg, sizeof(msg), "cannot open \"%.300s\" for writing", filename);
        stream_err(msg);
        return nullptr;
    }

    b200flac_stream* s = new b200flac_stream();
    s->f = f;
    s->params = *params;
    s->padding_size = padding_size;
    s->cur = 0; s->oldest = 0; s->n_in_flight = 0;
    s->next_frame_number = 0;
    s->total_samples = 0;
    s->min_frame = 0xFFFFFF; s->max_frame = 0; // flac.c:195-196
    s->current_offset = 0;
    s->failed = false;
    s->jobs_submitted = s->jobs_done = 0;
    s->md5_quit = false;
    md5_init(&s->md5);
    pthread_mutex_init(&s->mu, nullptr);
    pthread_cond_init(&s->cv_job, nullptr);
    pthread_cond_init(&s->cv_done, nullptr);
    pthread_create(&s->md5_thread, nullptr, md5_worker, s);

    // batch size: ~2048 blocks, at least one block, bounded to keep staging modest
    const uint32_t bs = params->block_size ? params->block_size : 1;
    uint64_t blocks = 2048;
    const uint64_t frame_bytes = (uint64_t)params->channels * (params->bits_per_sample / 8);
    while (blocks > 1 && blocks * bs * frame_bytes > (64ull << 20)) blocks /= 2;
    s->batch_frames = blocks * bs;

    // no device list (the Python entry point has no such parameter): device 0, or what B200FLAC_DEVICE names --
    // how one process per GPU (torchrun, a job scheduler) points the unchanged encode_flac call at its own GPU
    int dev0 = 0;
    if (!devices || n_devices <= 0) {
        const char* e = getenv("B200FLAC_DEVICE");
        if (e && *e) dev0 = atoi(e);
        devices = &dev0; n_devices = 1;
    }
    const int slots_per_dev = 2;
    s->slots_per_dev = slots_per_dev;
    s->reusable = true;
    for (int i = 0; i < n_devices; i++) {
        b200flac_encoder* e = pool_acquire(params, devices[i], s->batch_frames, slots_per_dev);
        if (!e) { stream_err(b200flac_last_error()); destroy_stream(s); return nullptr; }
        s->encs.push_back(e);
        s->enc_devices.push_back(devices[i]);
    }
    for (int k = 0; k < slots_per_dev; k++)
        for (int i = 0; i < n_devices; i++) {
            Lane l;
            l.enc = s->encs[i]; l.slot = k; l.pcm = b200flac_encoder_slot_pcm(l.enc, k);
            l.fill = 0; l.md5_enq = 0; l.seg_start = 0; l.in_flight = false; l.md5_ticket = 0; l.ext = nullptr;
            s->lanes.push_back(l);
        }

    // ---- stream head, flac.c:209-238 ----
    std::vector<uint8_t> head;
    const uint8_t zero_md5[16] = {0};
    b200flac_internal_stream_head(params, padding_size, version, s->min_frame, s->max_frame, s->total_samples, zero_md5, head);
    if (fwrite(head.data(), 1, head.size(), f) != head.size()) {
        stream_err("write error"); destroy_stream(s); return nullptr;
    }
    return s;
}

// collect the oldest in-flight lane and append its frames to the file
static int collect_oldest(b200flac_stream* s)
{
    Lane& l = s->lanes[s->oldest];
    // the frames land in the slot's pinned output buffer (a direct DMA) and are written to the file from there
    uint64_t out_cap = 0;
    uint8_t* out = b200flac_encoder_slot_out(l.enc, l.slot, &out_cap);
    if (!out) { stream_err(b200flac_last_error()); s->failed = true; return 1; }
    const size_t maxf = (size_t)(l.fill / s->params.block_size + l.segs.size() + 2);
    if (s->fbytes.size() < maxf) { s->fbytes.resize(maxf); s->fpcm.resize(maxf); }
    uint64_t nbytes = 0;
    uint32_t nfr = 0;
    if (b200flac_encoder_collect(l.enc, l.slot, out, out_cap, &nbytes, s->fbytes.data(),
                                 s->fpcm.data(), (uint32_t)maxf, &nfr)) {
        stream_err(b200flac_last_error());
        s->failed = true;
        return 1;
    }
    for (uint32_t i = 0; i < nfr; i++) {
        // flac.c:249-266
        s->offsets.push_back(s->current_offset);
        s->lengths.push_back(s->fpcm[i]);
        s->total_samples += s->fpcm[i];
        if (s->fbytes[i] < s->min_frame) s->min_frame = s->fbytes[i];
        if (s->fbytes[i] > s->max_frame) s->max_frame = s->fbytes[i];
        s->current_offset += s->fbytes[i];
    }
    if (nbytes && fwrite(out, 1, (size_t)nbytes, s->f) != nbytes) {
        stream_err("write error"); s->failed = true; return 1;
    }
    l.in_flight = false;
    l.fill = 0;
    l.md5_enq = 0;
    l.ext = nullptr;
    l.segs.clear();
    s->oldest = (s->oldest + 1) % s->lanes.size();
    s->n_in_flight--;
    return 0;
}

static void close_segment(b200flac_stream* s, Lane& l)
{
    if (l.fill > l.seg_start) {
        b200flac_segment g;
        g.pcm_frame_offset = l.seg_start;
        g.n_pcm_frames = l.fill - l.seg_start;
        g.first_frame_number = s->next_frame_number;
        g.reserved = 0;
        l.segs.push_back(g);
        const uint32_t bs = s->params.block_size;
        s->next_frame_number += (uint32_t)((g.n_pcm_frames + bs - 1) / bs);
        l.seg_start = l.fill;
    }
}

// Hands the staged PCM the MD5 thread has not seen yet to it -- while a batch is still filling, in pieces
// of at least `min_bytes`, so the (serial) hash runs under the copy or file read that fills the batch
// instead of starting when the batch is complete.  The lane's ticket is its last job.
static void md5_enqueue(b200flac_stream* s, Lane& l, size_t min_bytes)
{
    const size_t frame_bytes = (size_t)s->params.channels * (s->params.bits_per_sample / 8);
    const size_t pending = (size_t)(l.fill - l.md5_enq) * frame_bytes;
    if (pending == 0 || pending < min_bytes) return;
    pthread_mutex_lock(&s->mu);
    s->jobs.push_back(Md5Job{(l.ext ? l.ext : l.pcm) + (size_t)l.md5_enq * frame_bytes, pending});
    l.md5_ticket = ++s->jobs_submitted;
    pthread_cond_signal(&s->cv_job);
    pthread_mutex_unlock(&s->mu);
    l.md5_enq = l.fill;
}

// hand the current lane to its device and move on to the next lane
static int submit_current(b200flac_stream* s)
{
    Lane& l = s->lanes[s->cur];
    close_segment(s, l);
    if (l.segs.empty()) return 0;
    md5_enqueue(s, l, 0); // whatever of this batch the MD5 thread has not been given yet
    if (b200flac_encoder_submit(l.enc, l.slot, l.ext ? l.ext : l.pcm, l.segs.data(), (uint32_t)l.segs.size())) {
        stream_err(b200flac_last_error());
        s->failed = true;
        return 1;
    }
    l.in_flight = true;
    s->n_in_flight++;
    s->cur = (s->cur + 1) % s->lanes.size();
    Lane& nx = s->lanes[s->cur];
    if (nx.in_flight && collect_oldest(s)) return 1; // lanes are reused in order, so nx is the oldest
    md5_wait(s, nx.md5_ticket);                      // its staging must not be hashed any more
    nx.fill = 0; nx.md5_enq = 0; nx.seg_start = 0; nx.segs.clear(); nx.ext = nullptr;
    return 0;
}

extern "C" int b200flac_stream_write(b200flac_stream* s, const uint8_t* pcm, uint64_t n_pcm_frames)
{
    if (!s || s->failed) { if (!s) stream_err("stream is NULL"); return 1; }
    const size_t frame_bytes = (size_t)s->params.channels * (s->params.bits_per_sample / 8);
    while (n_pcm_frames) {
        Lane& l = s->lanes[s->cur];
        // a batch may only end on a block boundary of the open segment
        const uint32_t bs = s->params.block_size;
        const uint64_t cap = l.seg_start + (s->batch_frames - l.seg_start) / bs * bs;
        const uint64_t room = cap - l.fill;
        if (room == 0) { if (submit_current(s)) return 1; continue; }
        uint64_t take = n_pcm_frames < room ? n_pcm_frames : room;
        const uint64_t piece = MD5_PIECE / frame_bytes + 1; // copy in pieces so the MD5 thread starts early
        if (take > piece) take = piece;
        memcpy(l.pcm + (size_t)l.fill * frame_bytes, pcm, (size_t)take * frame_bytes);
        l.fill += take;
        md5_enqueue(s, l, MD5_PIECE);
        pcm += (size_t)take * frame_bytes;
        n_pcm_frames -= take;
        if (l.fill == cap && submit_current(s)) return 1;
    }
    return 0;
}

// file bytes -> signed little-endian, in place (FrameList(data, ..., is_big_endian, is_signed) of
// wav.py:523-527 / aiff.py:452-456 followed by the little-endian packing the MD5 callback sees)
static void to_signed_le(uint8_t* p, size_t n_samples, unsigned bytes, uint32_t flags)
{
    if ((flags & B200FLAC_PCM_BIG_ENDIAN) && bytes == 2) {
        uint16_t* w = (uint16_t*)p;
        for (size_t i = 0; i < n_samples; i++) w[i] = (uint16_t)((w[i] << 8) | (w[i] >> 8));
    } else if ((flags & B200FLAC_PCM_BIG_ENDIAN) && bytes == 3) {
        for (size_t i = 0; i < n_samples; i++, p += 3) { const uint8_t t = p[0]; p[0] = p[2]; p[2] = t; }
        p -= n_samples * 3;
    }
    if (flags & B200FLAC_PCM_UNSIGNED)
        for (size_t i = 0; i < n_samples; i++) p[i * bytes + bytes - 1] ^= 0x80; // value - 2^(bps-1)
}

extern "C" int b200flac_stream_write_file(b200flac_stream* s, const char* path, uint64_t byte_offset,
                                          uint64_t n_pcm_frames, uint32_t flags)
{
    if (!s || s->failed) { if (!s) stream_err("stream is NULL"); return 1; }
    if (!path) { stream_err("path is NULL"); return 1; }
    FILE* in = fopen(path, "rb");
    if (!in) {
        char msg[400];
        snprintf(msg, sizeof(msg), "cannot open \"%.300s\" for reading", path);
        stream_err(msg);
        return 1;
    }
    setvbuf(in, nullptr, _IONBF, 0); // whole batches go straight to the pinned staging
    if (fseeko(in, (off_t)byte_offset, SEEK_SET) != 0) { fclose(in); stream_err("seek error"); return 1; }
    const unsigned sample_bytes = s->params.bits_per_sample / 8;
    const size_t frame_bytes = (size_t)s->params.channels * sample_bytes;
    const uint32_t bs = s->params.block_size;
    int rc = 0;
    while (n_pcm_frames && !rc) {
        Lane& l = s->lanes[s->cur];
        const uint64_t cap = l.seg_start + (s->batch_frames - l.seg_start) / bs * bs;
        const uint64_t room = cap - l.fill;
        if (room == 0) { rc = submit_current(s); continue; }
        uint64_t take = n_pcm_frames < room ? n_pcm_frames : room;
        const uint64_t piece = MD5_PIECE / frame_bytes + 1; // read in pieces so the MD5 thread starts early
        if (take > piece) take = piece;
        uint8_t* dst = l.pcm + (size_t)l.fill * frame_bytes;
        const size_t want = (size_t)take * frame_bytes;
        if (fread(dst, 1, want, in) != want) {
            stream_err(ferror(in) ? "read error" : "premature end of data chunk");
            s->failed = true;
            rc = ferror(in) ? 1 : 2;
            break;
        }
        if (flags) to_signed_le(dst, (size_t)take * s->params.channels, sample_bytes, flags);
        l.fill += take;
        md5_enqueue(s, l, MD5_PIECE);
        n_pcm_frames -= take;
        if (l.fill == cap) rc = submit_current(s);
    }
    fclose(in);
    return rc;
}

// Force a frame boundary here: the PCM written since the last boundary that does not fill a
// whole block becomes a short frame (what the reference does when read() returns fewer frames
// than block_size mid-stream, flac.c:247,525).
extern "C" int b200flac_stream_end_block(b200flac_stream* s)
{
    if (!s || s->failed) return 1;
    Lane& l = s->lanes[s->cur];
    const uint32_t bs = s->params.block_size;
    if ((l.fill - l.seg_start) % bs == 0) return 0; // already on a boundary
    close_segment(s, l);
    // every short read adds a segment (and a frame) to the batch; an encoder has room for 1024 segment
    // tails per batch (b200flac_encoder_create), so a reader that keeps returning short reads ends the
    // batch early instead of overflowing it
    if (l.segs.size() >= 1000) return submit_current(s);
    return 0;
}

extern "C" int b200flac_stream_close(b200flac_stream* s, int abort_encode, uint64_t** frame_offsets,
                                     uint32_t** frame_pcm_frames, uint64_t* n_frames)
{
    if (!s) return 1;
    int rc = s->failed ? 1 : 0;
    if (!abort_encode && !rc) {
        if (s->lanes[s->cur].fill > 0 && submit_current(s)) rc = 1;
        while (!rc && s->n_in_flight) if (collect_oldest(s)) rc = 1;
    }
    if (!abort_encode && !rc) {
        md5_wait(s, s->jobs_submitted);
        uint8_t digest[16], si[34];
        md5_final(&s->md5, digest);
        build_streaminfo(s, digest, si);
        if (fseek(s->f, 8, SEEK_SET) != 0 || fwrite(si, 1, 34, s->f) != 34) { stream_err("write error"); rc = 1; }
        if (fflush(s->f) != 0) { stream_err("write error"); rc = 1; }
    }
    if (!rc && !abort_encode) {
        const size_t n = s->offsets.size();
        if (frame_offsets) {
            *frame_offsets = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
            memcpy(*frame_offsets, s->offsets.data(), n * sizeof(uint64_t));
        }
        if (frame_pcm_frames) {
            *frame_pcm_frames = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
            memcpy(*frame_pcm_frames, s->lengths.data(), n * sizeof(uint32_t));
        }
        if (n_frames) *n_frames = n;
    }
    destroy_stream(s);
    return rc;
}

extern "C" void b200flac_free(void* p) { free(p); }

// b200flac_stream_write for PCM in the caller's own page-locked memory that stays valid until the stream is closed
// (b200flac_encode_file): whole batches go to the device straight from it and the MD5 thread hashes it in place --
// no copy into the lanes' staging.  Only on a lane that holds nothing yet.
static int stream_write_in_place(b200flac_stream* s, const uint8_t* pcm, uint64_t n_pcm_frames)
{
    const size_t frame_bytes = (size_t)s->params.channels * (s->params.bits_per_sample / 8);
    const uint32_t bs = s->params.block_size;
    const uint64_t cap = s->batch_frames / bs * bs;
    while (n_pcm_frames) {
        Lane& l = s->lanes[s->cur];
        if (l.fill != 0 || l.seg_start != 0 || !l.segs.empty()) return b200flac_stream_write(s, pcm, n_pcm_frames);
        const uint64_t take = n_pcm_frames < cap ? n_pcm_frames : cap;
        l.ext = pcm;
        // the hash runs ahead in pieces, as it does under the staging copy
        const uint64_t piece = MD5_PIECE / frame_bytes + 1;
        while (l.fill < take) {
            l.fill += (take - l.fill) < piece ? (take - l.fill) : piece;
            md5_enqueue(s, l, 0);
        }
        if (submit_current(s)) return 1;
        pcm += (size_t)take * frame_bytes;
        n_pcm_frames -= take;
    }
    return 0;
}

extern "C" int b200flac_encode_file(const char* filename, const b200flac_params* params,
                                    uint32_t padding_size, const char* version,
                                    const uint8_t* pcm, uint64_t n_pcm_frames,
                                    const int* devices, int n_devices)
{
    b200flac_stream* s = b200flac_stream_open(filename, params, padding_size, version, devices, n_devices);
    if (!s) return 1;
    const int in_place = n_pcm_frames && b200flac_internal_is_pinned(pcm);
    if (in_place ? stream_write_in_place(s, pcm, n_pcm_frames) : b200flac_stream_write(s, pcm, n_pcm_frames)) {
        b200flac_stream_close(s, 1, nullptr, nullptr, nullptr);
        return 1;
    }
    return b200flac_stream_close(s, 0, nullptr, nullptr, nullptr);
}
