#!/bin/bash
# usage: tools/spills.sh <mangled-kernel-substring>   -- source lines of the spill loads/stores
# (STL/LDL) of a kernel in the freshly compiled push.cu (no GPU needed)
set -e
T=$(mktemp -d)
cd /root/repo/osiris_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -cubin push.cu -o $T/push.cubin
nvdisasm -g -c $T/push.cubin > $T/all.txt
awk -v pat="$1" '/^[ \t]*\.section[ \t]+\.text\./{infn=(index($0,pat)>0)} infn' $T/all.txt > $T/k.txt
echo "instructions: $(grep -c '^        /\*[0-9a-f]\{4\}\*/' $T/k.txt)"
awk '/\/\/## File/{f=$0} /STL|LDL/{print f}' $T/k.txt | sed -E 's/.*File "([^"]+)", line ([0-9]+).*/\1:\2/' | sed 's#.*/##' | sort | uniq -c | sort -rn | head -${2:-20}
