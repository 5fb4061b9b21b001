#!/bin/bash
# A/B builds of the library: scripts/build_variant.sh <name> "<-D flags>"  ->  variants/libuwbgo_<name>.so
# (select at run time with UWBGO_LIB=variants/libuwbgo_<name>.so)
set -e
cd "$(dirname "$0")/../localization_b200/csrc"
mkdir -p ../../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC $2"
nvcc $FLAGS -c uwbgo_kernels.cu -o /tmp/uwbgo_kernels_$1.o &
nvcc $FLAGS -c uwbgo_general_items.cu -o /tmp/uwbgo_general_items_$1.o &
nvcc $FLAGS -c uwbgo_window.cu -o /tmp/uwbgo_window_$1.o &
nvcc $FLAGS -c uwbgo_api.cu -o /tmp/uwbgo_api_$1.o &
nvcc $FLAGS -c uwbgo_stream.cu -o /tmp/uwbgo_stream_$1.o &
wait
nvcc $ARCH -shared -o ../../variants/libuwbgo_$1.so /tmp/uwbgo_kernels_$1.o /tmp/uwbgo_general_items_$1.o /tmp/uwbgo_window_$1.o /tmp/uwbgo_api_$1.o /tmp/uwbgo_stream_$1.o
echo built variants/libuwbgo_$1.so
