#!/bin/bash
# experiment build of the library: s2v_embed_tc.cu with extra -D flags, the other objects from the normal build
# usage: scripts/build_exp16.sh <out.so> [-D...]
set -e
cd "$(dirname "$0")/.."
out=$1; shift
python -m simmodel_b200.build > /dev/null
nvcc -c simmodel_b200/csrc/s2v_embed_tc.cu -o /tmp/s2v_embed_tc_exp.o -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -Xcompiler -O3 -I include -I simmodel_b200/csrc "$@"
objs=$(ls simmodel_b200/build/*.o | grep -v s2v_embed_tc.o)
nvcc -shared -o "$out" $objs /tmp/s2v_embed_tc_exp.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart
echo "$out"
