#!/bin/bash
# A/B tooling: builds libtrt_b200.so from a git revision (or the working tree, WORK) with extra nvcc flags (-D macros)
# into tools/_variants/libtrt_<name>.so without touching the in-tree build; tools/ab_variants.py times such libraries.
# usage: build_variant.sh <name> <git-rev|WORK> [extra nvcc flags...]
set -e
name=$1; rev=$2; shift 2
d=/tmp/var_$name; rm -rf $d; mkdir -p $d/pkg/csrc $d/include
if [ "$rev" = WORK ]; then cp -r /root/repo/tinyraytracing_b200/csrc/. $d/pkg/csrc/; cp /root/repo/include/*.h $d/include/
else git -C /root/repo archive $rev tinyraytracing_b200/csrc include | tar -x -C $d; mv $d/tinyraytracing_b200/csrc/* $d/pkg/csrc/; fi
rm -rf $d/pkg/csrc/build
mkdir -p $d/pkg/csrc/../../include; cp $d/include/*.h $d/pkg/csrc/../../include/ 2>/dev/null || true
cd $d/pkg/csrc
make -j8 ../libtrt_b200.so NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 -Xptxas -v $*" > $d/make.log 2>&1 || (tail -20 $d/make.log; exit 1)
cp ../libtrt_b200.so /root/repo/tools/_variants/libtrt_$name.so
grep -A2 "k_shadeE" build/wavefront.ptxas.log | grep "Used" | head -1
echo built $name
