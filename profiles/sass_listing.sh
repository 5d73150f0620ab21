#!/bin/sh
# Writes the SASS listings of the hot kernels (as built into python-audio-tools_b200/libb200flac.so) to
# profiles/r02_sass_<kernel>.txt:  cuobjdump -xelf + nvdisasm -c, one function per file.
#   sh profiles/sass_listing.sh
set -e
cd "$(dirname "$0")/.."
tmp=$(mktemp -d)
(cd "$tmp" && cuobjdump -xelf all "$OLDPWD/python-audio-tools_b200/libb200flac.so" > /dev/null)
one() {  # cubin, mangled-name pattern, output name
    nvdisasm -c "$tmp/$1" | awk -v pat="$2" '
        /^\.text\./ { on = ($0 ~ pat) }
        on { print }' > "profiles/r02_sass_$3.txt"
    echo "$3: $(grep -c "^ *//\*[0-9a-f]*\*/\|^ */\*[0-9a-f]*\*/" profiles/r02_sass_$3.txt) instructions"
}
one b200flac_encoder.sm_100a.cubin '_Z11k_lpc_autocILi12ELi1ELi0EE' k_lpc_autoc_12_1
one b200flac_encoder.sm_100a.cubin '_Z12k_analyze_v3ILi5ELb0ELi32ELi1ELb0EE' k_analyze_v3_5_0_32
one b200flac_encoder.sm_100a.cubin '_Z12k_analyze_v3ILi5ELb1ELi32ELi2ELb0EE' k_analyze_v3_exhaustive_order8
one b200flac_batch.sm_100a.cubin '_Z12k_md5_tracks' k_md5_tracks
one b200flac_encoder.sm_100a.cubin '_Z9k_pack_v3ILi256ELi4ELi32EE' k_pack_v3_256_4_32
one b200tta.sm_100a.cubin '_Z14k_tta_residualILj2EE' k_tta_residual_2
one b200alac.sm_100a.cubin '_Z11k_alac_size' k_alac_size
rm -rf "$tmp"
