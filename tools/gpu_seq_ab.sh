#!/bin/bash
# Same-box A/B of the sequential per-channel kernels (C4FM sync, voice-path MMSE): parity tests on the new library, then
# timings + output digests of the new and the previous build (ab_libs/lib_old.so), then the voice path's launch list.
O=gpurun_out/r03; mkdir -p $O
( time python -m pytest tests/test_c4fm_gpu.py tests/test_cqpsk_gpu.py tests/test_discriminator_gpu.py tests/test_p25_framer_gpu.py tests/test_fuzz_gpu.py -m gpu -q --timeout 900 ) > $O/seq_pytest.log 2>&1; echo "pytest rc=$?" >> $O/seq_pytest.log
tail -15 $O/seq_pytest.log
cp wavecap-sdr_b200/libwcsdr_b200.so /tmp/lib_new.so
rm -f $O/seq_ab.log
for v in new old new; do
  cp $( [ $v = new ] && echo /tmp/lib_new.so || echo ab_libs/lib_old.so ) wavecap-sdr_b200/libwcsdr_b200.so
  echo "== $v" | tee -a $O/seq_ab.log
  python tools/dev_c4fm.py 64 72000 5 2>&1 | tail -1 | tee -a $O/seq_ab.log
  python tools/dev_c4fm.py 1024 72000 3 2>&1 | tail -1 | tee -a $O/seq_ab.log
  python tools/dev_cqpsk.py 64 72000 5 2>&1 | tail -1 | tee -a $O/seq_ab.log
  python tools/dev_discdemod.py 64 72000 5 2>&1 | tail -1 | tee -a $O/seq_ab.log
  python tools/dev_discdemod.py 64 72000 5 aligned 2>&1 | tail -1 | tee -a $O/seq_ab.log
  python tools/dev_discdemod.py 1024 72000 3 2>&1 | tail -1 | tee -a $O/seq_ab.log
done
cp /tmp/lib_new.so wavecap-sdr_b200/libwcsdr_b200.so
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/discdemod_launches.csv python tools/dev_discdemod.py 64 72000 2 > $O/ncu_dd.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/c4fm_launches.csv python tools/dev_c4fm.py 64 72000 2 > $O/ncu_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/cqpsk_launches.csv python tools/dev_cqpsk.py 64 72000 2 > $O/ncu_cq.log 2>&1
