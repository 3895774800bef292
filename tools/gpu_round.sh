#!/bin/bash
# One GPU visit (1 GPU): parity tests, smoke, per-config benches, headline bench + reference arm, ncu launch lists,
# ncu --set full of the headline kernel and of the C2 kernels. Summaries: python tools/ncu_summary.py r<NN>_chan_fm --kernel chan256
# Multi-GPU legs (run separately, `gpurun --gpus N`):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus N
#   ... bench.py --gpus N --mode pull --steps 200        (ONE capture, kernels pull their slab over NVLink)
#   ... bench.py --gpus N --mode broadcast               (ONE capture, NCCL broadcast in front: the baseline)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python tools/bench_configs.py > gpurun_out/configs.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --chunks 16 --e2e-chunks 2 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chan256 -c 2 -o gpurun_out/prof_chan_fm -f \
    python bench.py --steps 2 --warmup 3 --chunks 16 --e2e-chunks 2 --no-cpu > gpurun_out/ncu2.log 2>&1
ncu -i gpurun_out/prof_chan_fm.ncu-rep --page raw --csv > gpurun_out/raw.csv 2>/dev/null
ncu -i gpurun_out/prof_chan_fm.ncu-rep --page source --csv > gpurun_out/src.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c2.csv \
    python tools/dev_c2only.py > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"front_kernel|resample_residue" -c 2 -o gpurun_out/prof_c2 -f \
    python tools/dev_c2only.py > gpurun_out/ncu_c2full.log 2>&1
ncu -i gpurun_out/prof_c2.ncu-rep --page raw --csv > gpurun_out/raw_c2.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; cat gpurun_out/bench.json
