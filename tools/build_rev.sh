#!/bin/bash
# Build the CUDA library of another revision next to the product build, for A/B runs on one box:
#   tools/build_rev.sh <git rev> <name>  ->  helen_b200/lib/libhelen_b200_<name>.so   (select with HB_LIB=...)
set -e
rev=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$rev" helen_b200/csrc include | tar -x -C "$tmp"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
     -o "$root/helen_b200/lib/libhelen_b200_$name.so" "$tmp/helen_b200/csrc/hb_api.cu"
rm -rf "$tmp"
echo "$root/helen_b200/lib/libhelen_b200_$name.so"
