#!/bin/bash
# quick GPU check: selected tests + a short bench without the CPU legs. usage: bash tools/gpu_quick.sh "<pytest args>" [bench args]
O=gpurun_out/r02
mkdir -p $O
( time python -m pytest $1 -m gpu -q --timeout 900 ) > $O/quick_pytest.log 2>&1; echo "pytest rc=$?" >> $O/quick_pytest.log
tail -25 $O/quick_pytest.log
if [ -n "$2" ]; then
  ( time python bench.py $2 ) > $O/quick_bench.json 2> $O/quick_bench.err; echo "bench rc=$?" >> $O/quick_bench.err
  tail -5 $O/quick_bench.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r02/quick_bench.json').read().strip().splitlines()[-1])
    print("value", l["value"], "frac", l["roofline"]["frac"], "sustained", (l.get("sustained") or {}).get("value"))
    print("e2e", l["e2e"]["value"], json.dumps(l["e2e"].get("modes"))[:1500])
except Exception as e:
    print("parse failed", e)
PY
fi
