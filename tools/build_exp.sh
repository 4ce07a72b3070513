#!/bin/bash
# Builds an experimental variant of the library next to the shipped one:  tools/build_exp.sh <n>  ->  svgf_b200/libsvgf_b200_exp<n>.so
# compiled with -DSVGF_EXP=<n> (see the #if SVGF_EXP blocks in svgf_b200/csrc).  A/B on the GPU box by copying it over
# libsvgf_b200.so between bench runs; nothing in the product selects it.
set -e
n=$1
here=$(cd "$(dirname "$0")/.." && pwd)
src=$here/svgf_b200/csrc
obj=$src/build_exp$n
mkdir -p $obj
tus="svgf_api svgf_tma svgf_band svgf_tu_packed_f16 svgf_tu_packed_f32 svgf_tu_staged_f16 svgf_tu_staged_f32 svgf_tu_lattice_f16 svgf_tu_lattice_f32 svgf_tu_stream svgf_tu_tiled svgf_tu_fused"
for t in $tus; do
  ( nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -DSVGF_EXP=$n -c -o $obj/$t.o $src/$t.cu 2> $obj/$t.ptxas.log || (cat $obj/$t.ptxas.log; exit 1) ) &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $here/svgf_b200/libsvgf_b200_exp$n.so $(for t in $tus; do echo $obj/$t.o; done) -ldl
echo built $here/svgf_b200/libsvgf_b200_exp$n.so
