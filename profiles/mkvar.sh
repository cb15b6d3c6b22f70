#!/bin/bash
# usage: profiles/mkvar.sh <name> [extra nvcc flags]  -> esloco_b200/_variants/<name>.so from the working tree
set -e
n=$1; shift
mkdir -p /root/repo/esloco_b200/_variants
cd /root/repo/esloco_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static "$@" -o ../_variants/$n.so esl_api.cu seqscan.cpp tablewriter.cpp -lz
echo built $n
