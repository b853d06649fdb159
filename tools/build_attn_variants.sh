#!/bin/bash
# Build libofab variants that differ only in attn.cu (register budget / pre-change source) for A/B timing on the GPU box.
set -e
cd "$(dirname "$0")/.."
python -m ofasys_b200.build >/dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
OBJS=$(ls ofasys_b200/build/*.o | grep -v attn.o)
mkdir -p ofasys_b200/variants
build() { # name, source, extra flags
  nvcc $FLAGS $3 -c $2 -o /tmp/attn_$1.o && nvcc -shared -o ofasys_b200/variants/libofab_$1.so $OBJS /tmp/attn_$1.o -gencode arch=compute_100a,code=sm_100a && echo built $1
}
git show ${BASE:-HEAD}:ofasys_b200/csrc/attn.cu > ofasys_b200/csrc/_attn_base.cu
build base ofasys_b200/csrc/_attn_base.cu "" &
build f4b3 ofasys_b200/csrc/attn.cu "" &
build f3b3 ofasys_b200/csrc/attn.cu "-DATTN_FWD_MINB=3" &
build f3b2 ofasys_b200/csrc/attn.cu "-DATTN_FWD_MINB=3 -DATTN_BWD_MINB=2" &
wait
rm -f ofasys_b200/csrc/_attn_base.cu
ls -la ofasys_b200/variants
