#!/bin/sh
# Developer tool: an A/B build of the engine library with extra compiler defines, next to the real one.
#   tools/build_variant.sh noroll -DV3_ROLL=0      ->  python-audio-tools_b200/libb200flac_noroll.so
#   B200FLAC_LIB=$PWD/python-audio-tools_b200/libb200flac_noroll.so python tools/pipe_probe.py 3600 0:3:1:1:0:0
set -e
cd "$(dirname "$0")/../python-audio-tools_b200"
tag=$1; shift
mkdir -p build/obj_$tag
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -fmad=false "$@" \
     -c -o build/obj_$tag/b200flac_encoder.o csrc/b200flac_encoder.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libb200flac_$tag.so build/obj_$tag/b200flac_encoder.o \
     build/obj/b200flac_stream.o build/obj/b200flac_batch.o build/obj/b200flac_metadata.o build/obj/b200flac_pcmfile.o build/obj/b200flac_decoder.o \
     build/obj/b200tta.o build/obj/b200alac.o build/obj/md5_lanes.o -lpthread
echo built libb200flac_$tag.so
