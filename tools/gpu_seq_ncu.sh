#!/bin/bash
# source-level ncu captures (stall samples per line) of the sequential per-channel kernels
O=gpurun_out/r03; mkdir -p $O
cap() {  # name regex command...
  n=$1; k=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/prof_$n -f "$@" > $O/ncu_$n.log 2>&1
  ncu -i $O/prof_$n.ncu-rep --page source --csv > $O/src_$n.csv 2>/dev/null
  ncu -i $O/prof_$n.ncu-rep --page raw --csv > $O/raw_$n.csv 2>/dev/null
  rm -f $O/prof_$n.ncu-rep
  tail -1 $O/ncu_$n.log
}
cap c4fm_sync c4fm_sync python tools/dev_c4fm.py 64 72000 1
cap dd_mmse dd_mmse python tools/dev_discdemod.py 64 72000 1

cap cqpsk_sync cqpsk_sync python tools/dev_cqpsk.py 64 72000 1
ls -la $O
