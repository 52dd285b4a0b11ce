#!/bin/bash
# ThreadSanitizer build of the library's host code + the ring harness (tests/native/ring_tsan.cpp) -> profiles/lab/bin/
set -e
cd "$(dirname "$0")/../.."
OUT=profiles/lab/bin
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=thread,-fno-omit-frame-pointer"
for u in b2s_runtime.cu b2s_columns.cu b2s_table.cu b2s_trees3.cu b2s_dense.cu; do
  nvcc $FLAGS -c mlrun_b200/csrc/$u -o $OUT/tsan_${u%.cu}.o &
done
nvcc $FLAGS -c mlrun_b200/csrc/b2s_codec.cpp -o $OUT/tsan_b2s_codec.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fsanitize=thread $OUT/tsan_*.o -o $OUT/libb200serve_tsan.so -Xlinker -ltsan
g++ -std=c++17 -O1 -g -fsanitize=thread tests/native/ring_tsan.cpp -o $OUT/ring_tsan -L$OUT -lb200serve_tsan -Wl,-rpath,'$ORIGIN' -L/usr/local/cuda/lib64 -lcudart -lpthread
ls -la $OUT/ring_tsan $OUT/libb200serve_tsan.so
