#!/bin/bash
# Builds a variant of the library for A/B timing: tools/build_variant.sh <name> <-D flags...>
# -> dynax_b200/_old/libdynax_<name>.so  (rc_api.cu and lti_chunked.cu recompiled with the flags, the rest reused)
set -e
cd "$(dirname "$0")/../dynax_b200"
name=$1; shift
mkdir -p _old _build/$name
for f in rc_api lti_chunked; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$f.cu -o _build/$name/$f.o &
done
wait
objs=$(ls _build/*.o | grep -v -e rc_api.o -e lti_chunked.o)
nvcc -shared -o _old/libdynax_$name.so $objs _build/$name/rc_api.o _build/$name/lti_chunked.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built _old/libdynax_$name.so
