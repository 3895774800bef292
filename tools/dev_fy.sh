for pb in 3 5; do for gy in 18 37 74 148; do WC_SPECTRUM_PASS_B=$pb WC_SPECTRUM_GY=$gy VALS=6 python tools/dev_spectrum.py | sed "s/^/pb=$pb gy=$gy /"; [ $pb = 3 ] && break; done; done
WC_SPECTRUM_PASS_B=5 python -m pytest tests/test_spectrum_gpu.py -x -q 2>&1 | tail -2
