#!/bin/bash
# Round-2 evidence visit (1 GPU): all -m gpu tests, smoke, the full bench line + reference arm, the reference's own
# benchmark_dsp.py through install(), ncu launch lists (bench, C1 plan, C2 plan, CQPSK) and the ncu --set full capture of the
# headline kernel. Everything lands in gpurun_out/ (tools/ncu_summary.py condenses it into profiles/).
O=gpurun_out
mkdir -p $O/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02/smi.txt 2>&1
( time python -m pytest tests -m gpu -q --timeout 900 ) > $O/r02/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02/smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02/smoke.log
( time python bench.py ) > $O/r02/bench.json 2> $O/r02/bench.err; echo "bench rc=$?" >> $O/r02/bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/r02/bench_ref.json 2>> $O/r02/bench.err
python - > $O/r02/reference_benchmark_dsp.log 2>&1 <<'PY'
import runpy, sys, os
sys.path.insert(0, os.getcwd())
from oracle import build_ref
if build_ref.staged():
    build_ref.load()
    import wavecap_sdr_b200.install as b200
    print("rebound:", len(b200.install(0)), "names")
    sys.argv = ["benchmark_dsp.py"]
    runpy.run_path(build_ref.benchmark_script(), run_name="__main__")
else:
    print("oracle/_ref not staged")
PY
B="python bench.py --steps 2 --warmup 3 --chunks 16 --e2e-chunks 2 --no-cpu --no-configs --no-modes --sustained-seconds 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $B > $O/r02/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chan256p -c 2 -o $O/prof_chan_fm -f $B > $O/r02/ncu2.log 2>&1
ncu -i $O/prof_chan_fm.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/prof_chan_fm.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
for c in c1 c2; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$c.csv python tools/dev_plan_only.py $c > $O/r02/ncu_$c.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_cqpsk.csv python tools/dev_cqpsk.py 64 72000 2 > $O/r02/ncu_cq.log 2>&1
tail -32 $O/r02/pytest_gpu.log; tail -1 $O/r02/smoke.log; tail -4 $O/r02/bench.err; tail -25 $O/r02/reference_benchmark_dsp.log
