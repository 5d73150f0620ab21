/* flac_oracle_main.c -- command-line twin of the reference's standalone
 * `flacenc` (src/encoders/flac.c:1637-1804): raw little-endian signed PCM on
 * stdin, same short options, FLAC file out.  TEST INFRASTRUCTURE ONLY.
 *
 * extra: --synth SEED:FRAMES generates the integer synthetic signal instead of
 * reading stdin; --dump-pcm FILE writes the PCM that was encoded. */
#include "flac_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void store_le(uint8_t *p, int32_t v, unsigned bytes)
{
    unsigned i;
    for (i = 0; i < bytes; i++) p[i] = (uint8_t)((uint32_t)v >> (8 * i));
}

int main(int argc, char **argv)
{
    orc_options opt;
    unsigned channels = 2, rate = 44100, bps = 16;
    const char *out = NULL, *dump = NULL;
    uint64_t synth_seed = 0, synth_frames = 0;
    int use_synth = 0, i;
    uint8_t *pcm = NULL, *file = NULL;
    size_t pcm_len = 0, cap = 0, file_len = 0, nfr = 0;
    FILE *f;

    memset(&opt, 0, sizeof(opt));
    opt.block_size = 4096; opt.max_lpc_order = 12; opt.max_residual_partition_order = 6;
    opt.padding_size = 4096;
    for (i = 1; i < argc; i++) {
        const char *a = argv[i];
#define NEXT (i + 1 < argc ? argv[++i] : "0")
        if (!strcmp(a, "-c")) channels = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "-r")) rate = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "-b")) bps = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "-B")) opt.block_size = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "-l")) opt.max_lpc_order = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "-P")) opt.min_residual_partition_order = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "-R")) opt.max_residual_partition_order = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "-m")) opt.mid_side = 1;
        else if (!strcmp(a, "-M")) opt.adaptive_mid_side = 1;
        else if (!strcmp(a, "-e")) opt.exhaustive_model_search = 1;
        else if (!strcmp(a, "--no-verbatim")) opt.no_verbatim_subframes = 1;
        else if (!strcmp(a, "--no-constant")) opt.no_constant_subframes = 1;
        else if (!strcmp(a, "--no-fixed")) opt.no_fixed_subframes = 1;
        else if (!strcmp(a, "--no-lpc")) opt.no_lpc_subframes = 1;
        else if (!strcmp(a, "--padding")) opt.padding_size = (unsigned)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "--synth")) {
            const char *v = NEXT;
            use_synth = 1;
            synth_seed = strtoull(v, NULL, 10);
            v = strchr(v, ':');
            synth_frames = v ? strtoull(v + 1, NULL, 10) : 0;
        } else if (!strcmp(a, "--dump-pcm")) dump = NEXT;
        else out = a;
    }
    if (!out) { fprintf(stderr, "usage: flacenc_oracle [options] out.flac < pcm\n"); return 1; }

    if (use_synth) {
        const unsigned bytes = bps / 8;
        const uint64_t chunk = 65536;
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * chunk * channels);
        uint64_t t;
        pcm_len = (size_t)(synth_frames * channels * bytes);
        pcm = (uint8_t *)malloc(pcm_len + 1);
        for (t = 0; t < synth_frames; t += chunk) {
            uint64_t n = synth_frames - t < chunk ? synth_frames - t : chunk, k;
            orc_synth_pcm(synth_seed, channels, bps, t, n, tmp);
            for (k = 0; k < n * channels; k++) store_le(pcm + (t * channels + k) * bytes, tmp[k], bytes);
        }
        free(tmp);
    } else {
        for (;;) {
            size_t got;
            if (cap - pcm_len < (1u << 20)) { cap = cap ? cap * 2 : (1u << 22); pcm = (uint8_t *)realloc(pcm, cap); }
            got = fread(pcm + pcm_len, 1, cap - pcm_len, stdin);
            if (got == 0) break;
            pcm_len += got;
        }
    }
    if (dump) { f = fopen(dump, "wb"); fwrite(pcm, 1, pcm_len, f); fclose(f); }
    orc_encode_stream(&opt, rate, channels, bps, pcm, pcm_len, &file, &file_len, NULL, NULL, &nfr);
    f = fopen(out, "wb");
    if (!f) { perror(out); return 1; }
    fwrite(file, 1, file_len, f);
    fclose(f);
    fprintf(stderr, "oracle: %zu frames, %zu bytes\n", nfr, file_len);
    orc_free(file);
    free(pcm);
    return 0;
}
