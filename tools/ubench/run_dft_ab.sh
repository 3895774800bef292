#!/bin/bash
# tcgen05-vs-FP32 DFT-256 A/B: timing run, then one ncu --set full capture per variant (profiles/r02_dft256_ab_*).
O=gpurun_out/r02
mkdir -p $O
timeout 120 tools/ubench/dft256_tc 4000 > $O/dft_ab.jsonl 2>&1; echo "rc=$?" >> $O/dft_ab.jsonl
for v in 0 1 2 3; do
  timeout 300 ncu --set full --clock-control none --import-source on -c 2 -o $O/prof_dft_v$v -f tools/ubench/dft256_tc 400 $v > $O/ncu_dft_v$v.log 2>&1
  ncu -i $O/prof_dft_v$v.ncu-rep --page raw --csv > $O/raw_dft_v$v.csv 2>/dev/null
done
cat $O/dft_ab.jsonl
