#!/bin/sh
# Builds build_trace/libvkgpu.so: the library with the per-phase HNSW hop trace compiled in (-DVKGPU_HNSW_TRACE).
# Run a binary against it with LD_LIBRARY_PATH=valkey_search_b200/csrc/build_trace (never the product build).
set -e
cd "$(dirname "$0")"
make -s -j8
mkdir -p build_trace
nvcc -gencode arch=compute_100a,code=sm_100a -DVKGPU_HNSW_TRACE -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr --extended-lambda -c hnsw.cu -o build_trace/hnsw.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_trace/libvkgpu.so build_trace/hnsw.o build/index.o build/flat_scan.o build/misc_kernels.o build/tensor_path.o build/batcher.o build/flat_select.o build/sharded.o -lcudart_static -lpthread -ldl -lrt
