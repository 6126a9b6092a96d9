timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python tools/late_profile.py c3o 300 2>&1 | tail -1 | cut -c1-300
