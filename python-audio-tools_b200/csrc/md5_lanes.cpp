// md5_lanes.cpp -- MD5 (RFC 1321) of up to sixteen byte strings at once on one host core.
//
// MD5 is a serial chain inside one string (a core does ~0.7 GB/s whatever its width), but strings are independent:
// sixteen chains in the lanes of one vector advance together at the cost of one.  b200flac_encode_files
// (b200flac_batch.cu) uses it for the host's share of the per-track STREAMINFO MD5s (flac.c:187-188, 277 hash every
// PCM byte of a stream): a pool thread takes up to sixteen tracks from the end of the list and hashes them side by
// side.  Written with GCC vector extensions and compiled three times (AVX-512: one 512-bit register per quantity,
// VPROLD / VPTERNLOGD; AVX2; baseline SSE2) -- the loader picks the clone for the CPU it runs on.  Plain g++, no CUDA.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define LANES 16
typedef uint32_t vu32 __attribute__((vector_size(4 * LANES)));

#define ROL(x, s) (((x) << (s)) | ((x) >> (32 - (s))))
#define F1(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define F2(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define F3(x, y, z) ((x) ^ (y) ^ (z))
#define F4(x, y, z) ((y) ^ ((x) | ~(z)))
#define STEP(f, w, x, y, z, data, k, s) (w += f(x, y, z) + (data) + (uint32_t)(k), w = ROL(w, s) + (x))

// the 64 steps on quantities a, b, c, d and message words x[16] of any type with +, ^, &, |, ~, <<, >> (one string: uint32_t;
// sixteen strings: a vector)
#define MD5_ROUNDS \
    STEP(F1, a, b, c, d, x[0], 0xd76aa478, 7);   STEP(F1, d, a, b, c, x[1], 0xe8c7b756, 12); \
    STEP(F1, c, d, a, b, x[2], 0x242070db, 17);  STEP(F1, b, c, d, a, x[3], 0xc1bdceee, 22); \
    STEP(F1, a, b, c, d, x[4], 0xf57c0faf, 7);   STEP(F1, d, a, b, c, x[5], 0x4787c62a, 12); \
    STEP(F1, c, d, a, b, x[6], 0xa8304613, 17);  STEP(F1, b, c, d, a, x[7], 0xfd469501, 22); \
    STEP(F1, a, b, c, d, x[8], 0x698098d8, 7);   STEP(F1, d, a, b, c, x[9], 0x8b44f7af, 12); \
    STEP(F1, c, d, a, b, x[10], 0xffff5bb1, 17); STEP(F1, b, c, d, a, x[11], 0x895cd7be, 22); \
    STEP(F1, a, b, c, d, x[12], 0x6b901122, 7);  STEP(F1, d, a, b, c, x[13], 0xfd987193, 12); \
    STEP(F1, c, d, a, b, x[14], 0xa679438e, 17); STEP(F1, b, c, d, a, x[15], 0x49b40821, 22); \
    STEP(F2, a, b, c, d, x[1], 0xf61e2562, 5);   STEP(F2, d, a, b, c, x[6], 0xc040b340, 9); \
    STEP(F2, c, d, a, b, x[11], 0x265e5a51, 14); STEP(F2, b, c, d, a, x[0], 0xe9b6c7aa, 20); \
    STEP(F2, a, b, c, d, x[5], 0xd62f105d, 5);   STEP(F2, d, a, b, c, x[10], 0x02441453, 9); \
    STEP(F2, c, d, a, b, x[15], 0xd8a1e681, 14); STEP(F2, b, c, d, a, x[4], 0xe7d3fbc8, 20); \
    STEP(F2, a, b, c, d, x[9], 0x21e1cde6, 5);   STEP(F2, d, a, b, c, x[14], 0xc33707d6, 9); \
    STEP(F2, c, d, a, b, x[3], 0xf4d50d87, 14);  STEP(F2, b, c, d, a, x[8], 0x455a14ed, 20); \
    STEP(F2, a, b, c, d, x[13], 0xa9e3e905, 5);  STEP(F2, d, a, b, c, x[2], 0xfcefa3f8, 9); \
    STEP(F2, c, d, a, b, x[7], 0x676f02d9, 14);  STEP(F2, b, c, d, a, x[12], 0x8d2a4c8a, 20); \
    STEP(F3, a, b, c, d, x[5], 0xfffa3942, 4);   STEP(F3, d, a, b, c, x[8], 0x8771f681, 11); \
    STEP(F3, c, d, a, b, x[11], 0x6d9d6122, 16); STEP(F3, b, c, d, a, x[14], 0xfde5380c, 23); \
    STEP(F3, a, b, c, d, x[1], 0xa4beea44, 4);   STEP(F3, d, a, b, c, x[4], 0x4bdecfa9, 11); \
    STEP(F3, c, d, a, b, x[7], 0xf6bb4b60, 16);  STEP(F3, b, c, d, a, x[10], 0xbebfbc70, 23); \
    STEP(F3, a, b, c, d, x[13], 0x289b7ec6, 4);  STEP(F3, d, a, b, c, x[0], 0xeaa127fa, 11); \
    STEP(F3, c, d, a, b, x[3], 0xd4ef3085, 16);  STEP(F3, b, c, d, a, x[6], 0x04881d05, 23); \
    STEP(F3, a, b, c, d, x[9], 0xd9d4d039, 4);   STEP(F3, d, a, b, c, x[12], 0xe6db99e5, 11); \
    STEP(F3, c, d, a, b, x[15], 0x1fa27cf8, 16); STEP(F3, b, c, d, a, x[2], 0xc4ac5665, 23); \
    STEP(F4, a, b, c, d, x[0], 0xf4292244, 6);   STEP(F4, d, a, b, c, x[7], 0x432aff97, 10); \
    STEP(F4, c, d, a, b, x[14], 0xab9423a7, 15); STEP(F4, b, c, d, a, x[5], 0xfc93a039, 21); \
    STEP(F4, a, b, c, d, x[12], 0x655b59c3, 6);  STEP(F4, d, a, b, c, x[3], 0x8f0ccc92, 10); \
    STEP(F4, c, d, a, b, x[10], 0xffeff47d, 15); STEP(F4, b, c, d, a, x[1], 0x85845dd1, 21); \
    STEP(F4, a, b, c, d, x[8], 0x6fa87e4f, 6);   STEP(F4, d, a, b, c, x[15], 0xfe2ce6e0, 10); \
    STEP(F4, c, d, a, b, x[6], 0xa3014314, 15);  STEP(F4, b, c, d, a, x[13], 0x4e0811a1, 21); \
    STEP(F4, a, b, c, d, x[4], 0xf7537e82, 6);   STEP(F4, d, a, b, c, x[11], 0xbd3af235, 10); \
    STEP(F4, c, d, a, b, x[2], 0x2ad7d2bb, 15);  STEP(F4, b, c, d, a, x[9], 0xeb86d391, 21);

// one string, scalar: what a lane falls back to when it is (nearly) alone -- one chain in a sixteen-lane vector runs at
// a third of the scalar code's speed
static void md5_scalar_blocks(uint32_t st[4], const uint8_t* p, size_t nblocks)
{
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    for (size_t blk = 0; blk < nblocks; blk++, p += 64) {
        uint32_t x[16];
        memcpy(x, p, 64);           // little-endian host
        const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
        MD5_ROUNDS
        a += a0; b += b0; c += c0; d += d0;
    }
    st[0] = a; st[1] = b; st[2] = c; st[3] = d;
}

// state[q][lane], q = a, b, c, d; lane l consumes `nblocks` 64-byte blocks at ptr[l], ptr[l] + stride[l], ...
// (stride 64, or 0 for a lane that has nothing to hash and rereads one dummy block)
__attribute__((target_clones("avx512f", "avx2", "default")))
static void md5_lanes_blocks(uint32_t state[4][LANES], const uint8_t* const ptr[LANES], const size_t stride[LANES], size_t nblocks)
{
    vu32 a, b, c, d;
    memcpy(&a, state[0], sizeof(a)); memcpy(&b, state[1], sizeof(b));
    memcpy(&c, state[2], sizeof(c)); memcpy(&d, state[3], sizeof(d));
    for (size_t blk = 0; blk < nblocks; blk++) {
        // word i of every lane's block (a 16 x 16 transpose of 32-bit words)
        vu32 x[16];
        uint32_t t[LANES][16];
        for (int l = 0; l < LANES; l++) memcpy(t[l], ptr[l] + stride[l] * blk, 64);
        for (int i = 0; i < 16; i++) {
            uint32_t col[LANES];
            for (int l = 0; l < LANES; l++) col[l] = t[l][i];
            memcpy(&x[i], col, sizeof(col));
        }
        const vu32 a0 = a, b0 = b, c0 = c, d0 = d;
        MD5_ROUNDS
        a += a0; b += b0; c += c0; d += d0;
    }
    memcpy(state[0], &a, sizeof(a)); memcpy(state[1], &b, sizeof(b));
    memcpy(state[2], &c, sizeof(c)); memcpy(state[3], &d, sizeof(d));
}

// Digests of n <= 16 byte strings.  `between` (may be NULL) is called every `piece_bytes` of progress per lane, so a
// pool thread can look after more urgent work; a non-zero return abandons the hashing (the digests are then unset).
extern "C" int b200flac_internal_md5_many(const uint8_t* const* ptr, const uint64_t* nbytes, uint32_t n, uint8_t* digests,
                                          uint64_t piece_bytes, int (*between)(void*), void* arg)
{
    if (n == 0) return 0;
    if (n > LANES) n = LANES;
    static const uint8_t nothing[64] = {0};
    uint32_t st[4][LANES];
    uint64_t done[LANES];
    for (int l = 0; l < LANES; l++) {
        st[0][l] = 0x67452301u; st[1][l] = 0xefcdab89u; st[2][l] = 0x98badcfeu; st[3][l] = 0x10325476u;
        done[l] = 0;
    }
    const uint64_t piece_blocks = piece_bytes >= 64 ? piece_bytes / 64 : 1;
    for (;;) {
        // all lanes with whole blocks left advance by the shortest remainder among them
        uint64_t step = 0;
        for (uint32_t l = 0; l < n; l++) {
            const uint64_t left = nbytes[l] / 64 - done[l];
            if (left && (!step || left < step)) step = left;
        }
        if (!step) break;
        if (step > piece_blocks) step = piece_blocks;
        uint32_t keep[4][LANES];
        memcpy(keep, st, sizeof(st));
        const uint8_t* p[LANES];
        size_t stride[LANES];
        bool idle[LANES];
        for (int l = 0; l < LANES; l++) {
            idle[l] = (uint32_t)l >= n || nbytes[l] / 64 == done[l];
            p[l] = idle[l] ? nothing : ptr[l] + 64 * done[l];
            stride[l] = idle[l] ? 0 : 64;
        }
        int active = 0;
        for (int l = 0; l < LANES; l++) active += idle[l] ? 0 : 1;
        if (active <= 2) {
            // (ragged lists: the long track that is left over when its neighbours are done)
            for (int l = 0; l < LANES; l++) {
                if (idle[l]) continue;
                uint32_t one[4] = {st[0][l], st[1][l], st[2][l], st[3][l]};
                md5_scalar_blocks(one, p[l], (size_t)step);
                for (int q = 0; q < 4; q++) st[q][l] = one[q];
                done[l] += step;
            }
        } else {
            md5_lanes_blocks(st, p, stride, (size_t)step);
            for (int l = 0; l < LANES; l++) {
                if (idle[l]) for (int q = 0; q < 4; q++) st[q][l] = keep[q][l];
                else done[l] += step;
            }
        }
        if (between && between(arg)) return 1;
    }
    // per string: the remaining bytes, 0x80, zeros, the length in bits -- one or two more blocks, all lanes again
    uint8_t tail[LANES][128];
    uint32_t tail_blocks[LANES];
    memset(tail, 0, sizeof(tail));
    for (int l = 0; l < LANES; l++) {
        tail_blocks[l] = 0;
        if ((uint32_t)l >= n) continue;
        const uint32_t rem = (uint32_t)(nbytes[l] & 63);
        if (rem) memcpy(tail[l], ptr[l] + (nbytes[l] - rem), rem);
        tail[l][rem] = 0x80;
        tail_blocks[l] = rem < 56 ? 1 : 2;
        const uint64_t bits = nbytes[l] << 3;
        for (int i = 0; i < 8; i++) tail[l][64 * tail_blocks[l] - 8 + i] = (uint8_t)(bits >> (8 * i));
    }
    for (uint32_t k = 0; k < 2; k++) {
        uint32_t keep[4][LANES];
        memcpy(keep, st, sizeof(st));
        const uint8_t* q[LANES];
        size_t stride[LANES];
        for (int l = 0; l < LANES; l++) { q[l] = tail[l] + 64 * k; stride[l] = 0; }
        md5_lanes_blocks(st, q, stride, 1);
        for (int l = 0; l < LANES; l++)
            if (k >= tail_blocks[l]) for (int w = 0; w < 4; w++) st[w][l] = keep[w][l];
    }
    for (uint32_t l = 0; l < n; l++)
        for (int i = 0; i < 16; i++) digests[16 * (size_t)l + i] = (uint8_t)(st[i >> 2][l] >> (8 * (i & 3)));
    return 0;
}
