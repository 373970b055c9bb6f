#!/bin/sh
# development build of the library with the attention event trace compiled in
cd "$(dirname "$0")/.." && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DDIP_TRACE -shared \
  diffusion-integer-programming_b200/csrc/*.cu -o tools/libdip_trace.so
