#!/bin/bash
# A/B of kernel-variant libraries on one box: tools/ab.sh <workload> <scale> <variant> [<variant> ...]  ("base" = the default library)
W=$1; S=$2; shift 2
for v in "$@"; do
  if [ "$v" = base ]; then unset OSGPU_LIB; else export OSGPU_LIB=$PWD/osiris_b200/libosiris_b200_$v.so; fi
  echo "== $v"; timeout 300 python tools/quickbench.py $W $S 2>&1 | tail -3 | head -2 | cut -c1-420
done
