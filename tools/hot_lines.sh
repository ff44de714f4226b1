#!/bin/bash
# hottest lines (stall samples) of the given kernels: tools/hot_lines.sh <out.txt> <view: sass|cuda> kernel-regex...
out=$1; view=$2; shift 2
mkdir -p /tmp/ncu
for k in "$@"; do
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -o /tmp/ncu/$k -f python tools/ncu_kernels.py 148 > /dev/null 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page source --csv --print-source $view > /tmp/ncu/$k.csv 2>/dev/null
  cp /tmp/ncu/$k.csv gpurun_out/hot_$k.csv; echo "=== $k"; python tools/ncu_hot_lines.py /tmp/ncu/$k.csv 45
done > $out 2>&1
