#!/bin/bash
# builds a tuning variant of the library: tools/build_variant.sh <name> <extra nvcc flags...>  -> natrium_b200/variants/lib_<name>.so
# (only the D3Q19 unit and the ABI unit are recompiled; the other stencil objects are reused from the regular build)
set -e
name=$1; shift
cd "$(dirname "$0")/../natrium_b200/csrc"
mkdir -p _build/var_$name ../variants
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $*"
nvcc $FL -c nb200.cu -o _build/var_$name/nb200.o 2> _build/var_$name/nb200.log &
nvcc $FL -DNB_D=3 -DNB_Q=19 -DNB_NAME=nb_ops_d3q19 -c inst.cu -o _build/var_$name/inst_d3q19.o 2> _build/var_$name/inst_d3q19.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so _build/var_$name/nb200.o _build/inst_d2q9.o _build/var_$name/inst_d3q19.o _build/inst_d3q15.o _build/inst_d2q25.o _build/inst_d3q45.o -ldl
grep -A2 "k_stream_collide_f_gridILi3ELi19ELi0" _build/var_$name/inst_d3q19.log | tail -2
