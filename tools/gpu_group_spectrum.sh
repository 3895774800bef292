#!/bin/bash
# dev: group-kernel spectrum vs the two-pass pipeline (WC_DEV library in ab_libs/), plus the DRAM bytes of one launch of each
cp wavecap-sdr_b200/libwcsdr_b200.so /tmp/keep.so
cp ab_libs/lib_dev_group.so wavecap-sdr_b200/libwcsdr_b200.so
mkdir -p gpurun_out/r02
timeout 300 python tools/dev_spectrum.py 2>&1 | tail -8
cat > /tmp/one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from wavecap_sdr_b200.dsp.fft.cuda_backend import CudaFFTBackend
be = CudaFFTBackend(65536)
x = torch.view_as_complex(torch.randn((1024 * 65536, 2), device="cuda") * 0.2)
for _ in range(3):
    be.execute_frames(x, 1024, 65536, 4)
torch.cuda.synchronize()
PY
for v in 3 6; do
  WC_SPECTRUM_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio --clock-control none -k regex:spectrum --csv --log-file gpurun_out/r02/ncu_group_v$v.csv python /tmp/one.py > /dev/null 2>&1
  echo "variant $v"; tail -24 gpurun_out/r02/ncu_group_v$v.csv | cut -d, -f5,13- | tail -16
done
cp /tmp/keep.so wavecap-sdr_b200/libwcsdr_b200.so
