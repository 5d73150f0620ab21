// b200flacenc -- the standalone driver of the reference encoder (src/encoders/flac.c:1637-1804,
// `flacenc [options] <output.flac> < raw PCM`) on the B200 engine: same options, same defaults, same
// file.  Raw PCM on stdin is little-endian signed at bits-per-sample/8 bytes per sample, interleaved
// (flac.c:1784-1790).  Links libb200flac.so (stream layer); no CPU fallback.
#include <errno.h>
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/b200flac.h"

static int usage(void)
{
    puts("*** Usage: b200flacenc [options] <output.flac>");
    puts("-c, --channels=#          number of input channels");
    puts("-r, --sample_rate=#       input sample rate in Hz");
    puts("-b, --bits-per-sample=#   bits per input sample");
    puts("");
    puts("-B, --block-size=#              block size");
    puts("-l, --max-lpc-order=#           maximum LPC order");
    puts("-P, --min-partition-order=#     minimum partition order");
    puts("-R, --max-partition-order=#     maximum partition order");
    puts("-m, --mid-side                  use mid-side encoding");
    puts("-M, --adaptive-mid-side         use adaptive mid-side encoding");
    puts("-e, --exhaustive-model-search   search for best subframe exhaustively");
    puts("-q, --quiet                     do not print the parameters");
    return 0;
}

static bool parse_uint(const char* s, const char* what, unsigned* out)
{
    char* end = nullptr;
    errno = 0;
    const unsigned long v = strtoul(s, &end, 10);
    if (errno || end == s || *end) { printf("invalid --%s \"%s\"\n", what, s); return false; }
    *out = (unsigned)v;
    return true;
}

int main(int argc, char* argv[])
{
    // defaults of the reference's driver, flac.c:1639-1650
    b200flac_params p;
    memset(&p, 0, sizeof(p));
    p.channels = 2; p.sample_rate = 44100; p.bits_per_sample = 16;
    p.block_size = 4096; p.max_lpc_order = 12;
    p.min_residual_partition_order = 0; p.max_residual_partition_order = 6;
    const char* output = nullptr;
    bool quiet = false;

    static const struct option long_opts[] = {
        {"help", no_argument, nullptr, 'h'},
        {"channels", required_argument, nullptr, 'c'},
        {"sample-rate", required_argument, nullptr, 'r'},
        {"bits-per-sample", required_argument, nullptr, 'b'},
        {"block-size", required_argument, nullptr, 'B'},
        {"max-lpc-order", required_argument, nullptr, 'l'},
        {"min-partition-order", required_argument, nullptr, 'P'},
        {"max-partition-order", required_argument, nullptr, 'R'},
        {"mid-side", no_argument, nullptr, 'm'},
        {"adaptive-mid-side", no_argument, nullptr, 'M'},
        {"exhaustive-model-search", no_argument, nullptr, 'e'},
        {"quiet", no_argument, nullptr, 'q'},
        {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "-hc:r:b:B:l:P:R:mMeq", long_opts, nullptr)) != -1) {
        switch (c) {
        case 1:
            if (output) { puts("only one output file allowed"); return 1; }
            output = optarg;
            break;
        case 'c': if (!parse_uint(optarg, "channel", &p.channels)) return 1; break;
        case 'r': if (!parse_uint(optarg, "sample-rate", &p.sample_rate)) return 1; break;
        case 'b': if (!parse_uint(optarg, "bits-per-sample", &p.bits_per_sample)) return 1; break;
        case 'B': if (!parse_uint(optarg, "block-size", &p.block_size)) return 1; break;
        case 'l': if (!parse_uint(optarg, "max-lpc-order", &p.max_lpc_order)) return 1; break;
        case 'P': if (!parse_uint(optarg, "min-partition-order", &p.min_residual_partition_order)) return 1; break;
        case 'R': if (!parse_uint(optarg, "max-partition-order", &p.max_residual_partition_order)) return 1; break;
        case 'm': p.mid_side = 1; break;
        case 'M': p.adaptive_mid_side = 1; break;
        case 'e': p.exhaustive_model_search = 1; break;
        case 'q': quiet = true; break;
        default: return usage();
        }
    }
    if (!output) { puts("exactly 1 output file required"); return 1; }
    if (p.channels < 1 || p.channels > 8 || (p.bits_per_sample != 8 && p.bits_per_sample != 16 && p.bits_per_sample != 24) ||
        p.sample_rate == 0) {
        puts("channels must be 1..8, bits per sample 8, 16 or 24, sample rate positive");
        return 1;
    }
    if (!quiet) {
        printf("Encoding from stdin using parameters:\n");
        printf("channels        %u\nsample rate     %u\nbits per sample %u\nlittle-endian, signed samples\n\n",
               p.channels, p.sample_rate, p.bits_per_sample);
        printf("block size              %u\nmax LPC order           %u\nmin partition order     %u\n"
               "max partition order     %u\nmid side                %d\nadaptive mid side       %d\n"
               "exhaustive model search %d\n",
               p.block_size, p.max_lpc_order, p.min_residual_partition_order, p.max_residual_partition_order,
               p.mid_side, p.adaptive_mid_side, p.exhaustive_model_search);
    }

    b200flac_stream* s = b200flac_stream_open(output, &p, 4096, nullptr, nullptr, 0);   // DEFAULT_PADDING_SIZE, flac.c:33
    if (!s) { fprintf(stderr, "*** Error encoding FLAC file \"%s\": %s\n", output, b200flac_last_error()); return 1; }
    const size_t frame_bytes = (size_t)p.channels * (p.bits_per_sample / 8);
    std::vector<uint8_t> buf((4u << 20) / frame_bytes * frame_bytes + frame_bytes);
    size_t have = 0;
    int rc = 0;
    for (;;) {
        const size_t r = fread(buf.data() + have, 1, buf.size() - have, stdin);
        have += r;
        const size_t frames = have / frame_bytes;
        if (frames && b200flac_stream_write(s, buf.data(), frames)) { rc = 1; break; }
        const size_t used = frames * frame_bytes;
        memmove(buf.data(), buf.data() + used, have - used);   // a partial PCM frame waits for its other bytes
        have -= used;
        if (r == 0) break;                                      // (a trailing partial PCM frame is dropped, as fread-by-frame does)
    }
    if (b200flac_stream_close(s, rc, nullptr, nullptr, nullptr)) rc = 1;
    if (rc) { fprintf(stderr, "*** Error encoding FLAC file \"%s\": %s\n", output, b200flac_last_error()); return 1; }
    return 0;
}
