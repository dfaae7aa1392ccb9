#!/bin/bash
# Build experiment variants of the library (C3 shape only) into build_variants/<name>.so.
# usage: scripts/variants.sh name1:patchcmd1 name2:patchcmd2 ...   (patchcmd runs inside the copied csrc dir)
set -e
ROOT=$(cd $(dirname $0)/.. && pwd)
for spec in "$@"; do
  name=${spec%%:*}; cmd=${spec#*:}
  d=$ROOT/build_variants/src_$name
  rm -rf $d; mkdir -p $d/daqp_b200 $d/include
  cp -r $ROOT/daqp_b200/csrc $d/daqp_b200/; cp $ROOT/include/daqp_b200.h $d/include/
  (cd $d/daqp_b200/csrc && eval "$cmd")
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DDAQP_B200_FAST_BUILD \
       -o $ROOT/build_variants/$name.so $d/daqp_b200/csrc/daqp_b200.cu &
done
wait
ls -la $ROOT/build_variants/*.so
