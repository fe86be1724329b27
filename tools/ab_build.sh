#!/bin/bash
# Build a named variant of libvit_b200.so for a same-box A/B (boxes of the pool differ by up to 10 %, so two
# builds are only comparable when they alternate inside ONE gpurun call -- profiles/r01_v6_attention.md).
#   tools/ab_build.sh NAME ["-DVITCU_GELU_FORM=0 ..."]     -> vit-with-opencl_b200/build/ab/NAME.so
# then:  gpurun -- 'tools/ab_run.sh 3 A B -- python tools/gemm_bench.py 256 20'
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; shift
mkdir -p "$root/vit-with-opencl_b200/build/ab"
touch "$root"/vit-with-opencl_b200/csrc/*.cu
make -C "$root/vit-with-opencl_b200" -j8 EXTRA="$*" > /dev/null
cp "$root/vit-with-opencl_b200/libvit_b200.so" "$root/vit-with-opencl_b200/build/ab/$name.so"
echo "built build/ab/$name.so with EXTRA='$*' (rebuild the default with 'make -C vit-with-opencl_b200' when done)"
