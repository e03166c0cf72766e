#!/bin/bash
# tools/build_at.sh <git-ref> <out.so>: build the library from the sources of a given commit (A/B timing on one box:
# ZEDO_B200_LIB=<out.so> python tools/layer_bench.py ...)
set -e
ref=$1; out=$2
tmp=$(mktemp -d)
git archive "$ref" zedo_release_b200/csrc include | tar -x -C "$tmp"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o "$out" "$tmp"/zedo_release_b200/csrc/*.cu
rm -rf "$tmp"
echo "built $out from $ref"
