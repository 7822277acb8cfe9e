#!/bin/bash
# round 2, GPU call 8: barotropic substep kernel variants (parity + microbench), streaming PLM edge values (parity + timing)
mkdir -p gpurun_out
( MOM6CU_BT_OPT=3 timeout 300 python -m pytest tests/test_bt_timeloop_gpu.py tests/test_btstep.py -m gpu -x -q > gpurun_out/r02_bt_opt3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bt_opt3.log )
tail -3 gpurun_out/r02_bt_opt3.log
( timeout 300 python -m pytest tests/test_pressure_force.py -m gpu -x -q > gpurun_out/r02_pgf_stream.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pgf_stream.log )
tail -3 gpurun_out/r02_pgf_stream.log
for o in 0 1 2 3; do ( MOM6CU_BT_OPT=$o timeout 200 python tools/bt_microbench.py 2>&1 | tail -1 ) >> gpurun_out/r02_bt_microbench.log; done
cat gpurun_out/r02_bt_microbench.log
( MOM6CU_PGF_RECON=1 timeout 300 python tools/prof_stage.py pgf 1440 1080 75 3 2>&1 | tail -2 ) > gpurun_out/r02_pgf_time_v3.log; cat gpurun_out/r02_pgf_time_v3.log
