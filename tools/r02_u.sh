set -x
PB200_DEBUG_CHECK=1 timeout 200 python tools/steps_diag.py c3o > gpurun_out/steps_diag_c3o_r02u.log 2>&1; grep -E "steps:|check:" gpurun_out/steps_diag_c3o_r02u.log | cut -c1-330 | head -40
