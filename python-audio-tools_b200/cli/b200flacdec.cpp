// b200flacdec -- the standalone driver of the reference decoder (src/decoders/flac.c:1340-1527,
// `flacdec <file.flac> > raw PCM`) on the B200 engine's frame-parallel decoder: PCM to stdout as signed
// little-endian, the reference's messages on stderr, exit status 1 on any error.  The whole stream is
// decoded at once, so a stream with a damaged frame produces no PCM at all, where the reference has
// already written the frames that precede it; an MD5 mismatch is reported after the PCM, as there.
// Links libb200flac.so; no CPU fallback.
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/b200flac.h"

int main(int argc, char* argv[])
{
    if (argc < 2) {
        fprintf(stderr, "*** Usage: %s <file.flac>\n", argv[0]);
        return 1;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) {
        fprintf(stderr, "*** %s: %s\n", argv[1], strerror(errno));
        return 1;
    }
    fseeko(f, 0, SEEK_END);
    const uint64_t n = (uint64_t)ftello(f);
    fseeko(f, 0, SEEK_SET);
    std::vector<uint8_t> data(n + 1);
    const size_t got = fread(data.data(), 1, n, f);
    fclose(f);
    b200flac_stream_info info;
    if (got != n || b200flac_read_streaminfo(data.data(), n, &info)) {
        fprintf(stderr, "*** Error reading streaminfo\n");
        return 1;
    }
    const uint64_t pcm_bytes = info.total_pcm_frames * info.channels * (info.bits_per_sample / 8);
    std::vector<uint8_t> pcm(pcm_bytes + 1);
    const int rc = b200flac_decode_memory(data.data(), n, 0, pcm.data(), pcm_bytes, nullptr, 1, nullptr, nullptr, nullptr,
                                          nullptr);
    const char* msg = b200flac_last_error();
    const bool md5_mismatch = rc == 1 && strcmp(msg, "MD5 mismatch at end of stream") == 0;
    if (rc == 0 || md5_mismatch) {
        if (pcm_bytes && fwrite(pcm.data(), 1, pcm_bytes, stdout) != pcm_bytes) {
            fprintf(stderr, "*** %s\n", strerror(errno));
            return 1;
        }
        fflush(stdout);
    }
    if (md5_mismatch) {
        fprintf(stderr, "*** MD5 mismatch at end of stream\n");         // flac.c:1499
        return 1;
    }
    if (rc == 2) {
        fprintf(stderr, "*** I/O Error reading frame\n");               // flac.c:1470
        return 1;
    }
    if (rc) {
        fprintf(stderr, "*** Error: %s\n", msg);                        // flac.c:1424, :1445, :1459
        return 1;
    }
    return 0;
}
