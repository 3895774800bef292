#!/bin/bash
# Same-box A/B of prebuilt library variants in ab_libs/ (git-ignored): each variant's short bench, ROUNDS times, interleaved.
# usage: bash tools/ab_libs.sh "old fft fold fft_fold" [rounds]
O=gpurun_out/r02; mkdir -p $O
cp wavecap-sdr_b200/libwcsdr_b200.so /tmp/lib_keep.so
R=${2:-2}
for r in $(seq 1 $R); do
  for v in $1; do
    cp ab_libs/lib_$v.so wavecap-sdr_b200/libwcsdr_b200.so
    python bench.py --no-cpu --no-one-capture --no-modes --sustained-seconds 1 > $O/ab_$v.$r.json 2> $O/ab_$v.$r.err
    python - $v $r <<'PY'
import json, sys
v, r = sys.argv[1], sys.argv[2]
try:
    l = json.loads(open(f"gpurun_out/r02/ab_{v}.{r}.json").read().strip().splitlines()[-1])
    c = {x["config"]: x["value"] for x in l.get("configs", [])}
    print(f"{v:10s} round {r}: fm {l['value']:.0f}  sustained {(l.get('sustained') or {}).get('value')}  mode0 {l['roofline']['channelizer_only']['msps_per_gpu']:.0f}  C3 {c.get('C3')}  C1 {c.get('C1')} C2 {c.get('C2')}  clk {l['clocks']['sm_mhz']}")
except Exception as e:
    print(v, r, "failed", e)
PY
  done
done
cp /tmp/lib_keep.so wavecap-sdr_b200/libwcsdr_b200.so
