#!/bin/bash
# Round-2 GPU visit (1 GPU): every -m gpu test (no -x: all failures in one visit), smoke, the full bench line (C5 burst +
# sustained, e2e + copy ceiling, configs C1-C4 with CPU arms) and the reference arm. Outputs under gpurun_out/r02/.
O=gpurun_out/r03
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt; lscpu | head -30 >> $O/nproc.txt; numactl -H >> $O/nproc.txt 2>&1
( time python -m pytest tests -m gpu -q --timeout 900 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time python bench.py ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.json 2>> $O/bench.err
tail -40 $O/pytest_gpu.log; tail -2 $O/smoke.log; tail -5 $O/bench.err; cut -c1-1500 $O/bench.json
for t in "dev_c4fm.py 64 72000 5" "dev_cqpsk.py 64 72000 5" "dev_discdemod.py 64 72000 5" "dev_discdemod.py 1024 72000 3"; do python tools/$t 2>&1 | tail -1 | tee -a $O/seq_now.log; done
