set -x
PB200_DEBUG_CHECK=1 timeout 900 python -m pytest "tests/test_gpu_parity.py::test_bucket_sort_overflow_and_skew_fall_back" -m gpu -q -x -s 2>&1 | grep -E "check:|Error|assert|crowd" | tail -30 | cut -c1-330
