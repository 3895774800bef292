#!/bin/bash
# ncu --set full of the headline kernel (2 launches) + launch list; summaries via tools/ncu_summary.py afterwards
O=gpurun_out; mkdir -p $O/r02
B="python bench.py --steps 2 --warmup 3 --chunks 16 --e2e-chunks 2 --no-cpu --no-configs --no-modes --no-one-capture --sustained-seconds 0"
$B > $O/r02/ncu_pre.json 2> $O/r02/ncu_pre.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $B > $O/r02/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chan256p -c 2 -o $O/prof_chan_fm -f $B > $O/r02/ncu2.log 2>&1
ncu -i $O/prof_chan_fm.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/prof_chan_fm.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
tail -2 $O/r02/ncu2.log
