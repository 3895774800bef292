#!/bin/bash
# dev: analog parity tests, C1/C2 step times and the front kernel's time under the launch list
mkdir -p gpurun_out
python -m pytest tests/test_analog_gpu.py -m gpu -q --timeout 600 2>&1 | tail -1
python tools/bench_configs.py 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try:
        d=json.loads(line)
    except Exception: continue
    if isinstance(d, dict) and d.get('config') in ('C1','C2'): print(d['config'], d['ms'], d['value'], d.get('python_batch_api_ms'), d.get('stage_by_stage_path_ms'))
"
for c in c1 c2; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$c.csv python tools/dev_plan_only.py $c > /dev/null 2>&1
grep front_kernel gpurun_out/launches_$c.csv | tail -1 | cut -d, -f8,9,15-
done
