O=gpurun_out/r03; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/c4fm_launches.csv python tools/dev_c4fm.py 64 72000 2 > $O/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:c4fm_sync -s 1 -c 1 -o $O/prof_c4 -f python tools/dev_c4fm.py 64 72000 1 > $O/ncu_c4s.log 2>&1
ncu -i $O/prof_c4.ncu-rep --page source --csv > $O/src_c4fm_sync.csv 2>/dev/null
ncu -i $O/prof_c4.ncu-rep --page raw --csv > $O/raw_c4fm_sync.csv 2>/dev/null
rm -f $O/prof_c4.ncu-rep
