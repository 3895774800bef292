#!/bin/bash
# One GPU visit: parity tests, dev sweep, bench, ncu launch list, ncu --set full of the headline kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
RS=32,64,128 python tools/dev_chan.py > gpurun_out/dev_chan.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --chunks 16 --e2e-chunks 2 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chan256 -c 2 -o gpurun_out/prof_chan_fm -f \
    python bench.py --steps 2 --warmup 3 --chunks 16 --e2e-chunks 2 --no-cpu > gpurun_out/ncu2.log 2>&1
ncu -i gpurun_out/prof_chan_fm.ncu-rep --page raw --csv > gpurun_out/raw.csv 2>/dev/null
ncu -i gpurun_out/prof_chan_fm.ncu-rep --page source --csv > gpurun_out/src.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
