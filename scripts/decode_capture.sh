#!/bin/bash
# GPU box: ncu --set full of the three decode kernels of the current build (one launch each, 1024 x 60 s streams without a block index).
# usage: scripts/decode_capture.sh <tag>   -> gpurun_out/<tag>_decode_{transform,parse,walk}_ncu_summary.txt (reports stay on the box)
tag=$1; out=gpurun_out; mkdir -p $out
for k in transform parse walk; do
  ncu --set full --clock-control none --import-source on -k regex:bp_${k}_kernel -c 1 -o /tmp/${tag}_$k \
      python scripts/time_decode.py --streams 1024 --mode walk --distinct 256 --reps 1 > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/${tag}_$k.ncu-rep --lines 30 > $out/${tag}_decode_${k}_ncu_summary.txt 2>&1
  head -12 $out/${tag}_decode_${k}_ncu_summary.txt | grep -E "kernel|duration|issue_active|dram__bytes"
done
