#!/bin/bash
# Build a variant of the library with extra preprocessor flags:  scripts/build_variant.sh <name> <flags...>
# -> simmodel_b200/libs2v_<name>.so  (used by the timeline / ablation scripts; never loaded by the product path)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p simmodel_b200/build/$name
objs=""
for f in simmodel_b200/csrc/*.cu; do
  o=simmodel_b200/build/$name/$(basename ${f%.cu}).o
  rm -f $o
  nvcc -c $f -o $o -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I include -I simmodel_b200/csrc "$@" &
  objs="$objs $o"
done
wait
for o in $objs; do [ -f $o ] || { echo "compile failed: $o" >&2; exit 1; }; done
nvcc -shared -o simmodel_b200/libs2v_$name.so $objs -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart
echo simmodel_b200/libs2v_$name.so
