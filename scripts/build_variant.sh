#!/bin/bash
# Developer helper: build libpcc_search with extra -D flags into _variants/<name>.so (git-ignored; travels with gpurun).
# usage: scripts/build_variant.sh name "-DPCC_X_..."      then: PCC_SO=_variants/name.so python scripts/probe.py ...
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
d=$root/_variants/$name
mkdir -p $d/pointcloudcomparator_b200
rm -rf $d/pointcloudcomparator_b200/csrc
cp -r $root/pointcloudcomparator_b200/csrc $d/pointcloudcomparator_b200/csrc
rm -f $d/pointcloudcomparator_b200/csrc/*.o
ln -sfn $root/include $d/include
make -s -C $d/pointcloudcomparator_b200/csrc -j6 EXTRA="$*" OUT=$root/_variants/$name.so
grep -A2 "knn_thr_kernelILi16ELb1" $d/pointcloudcomparator_b200/csrc/pcc_knn.ptxas.log | grep -E -o "Used [0-9]* registers|[0-9]* bytes spill stores" | head -2
rm -rf $d
