#!/bin/bash
# Builds a variant of libgknext_cuda.so for same-box A/B runs: scripts/build_variant.sh NAME [nvcc -D flags...]
# -> build_ab/NAME/libgknext_cuda.so (select with GKCU_LIB=...). build_ab/ is git-ignored but travels to the GPU box.
set -e
N=$1; shift
D=build_ab/$N
rm -rf $D && mkdir -p $D/gknextengine_b200 $D/include
cp -r gknextengine_b200/csrc $D/gknextengine_b200/ && cp include/*.h $D/include/
rm -f $D/gknextengine_b200/csrc/*.o
make -s -C $D/gknextengine_b200/csrc -j4 EXTRA="$*" >/dev/null
mv $D/gknextengine_b200/libgknext_cuda.so $D/ && rm -rf $D/gknextengine_b200 $D/include
ls -la $D/libgknext_cuda.so
