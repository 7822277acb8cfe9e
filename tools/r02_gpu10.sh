#!/bin/bash
# round 2, GPU call 10: A/B of the barotropic substep tile and of the PressureForce reconstruction kernel variants (parity for the ones kept)
mkdir -p gpurun_out
for t in 0 1 2; do ( MOM6CU_BT_TILE=$t timeout 200 python tools/bt_microbench.py 2>&1 | tail -1 | sed "s/^/TILE=$t /" ) >> gpurun_out/r02_bt_tiles.log; done
cat gpurun_out/r02_bt_tiles.log
for v in 0 1 2 3; do ( MOM6CU_PGF_VAR=$v MOM6CU_PGF_RECON=1 timeout 300 python tools/prof_stage.py pgf 1440 1080 75 3 2>&1 | tail -1 | sed "s/^/VAR=$v /" ) >> gpurun_out/r02_pgf_vars.log; done
cat gpurun_out/r02_pgf_vars.log
( MOM6CU_BT_TILE=1 MOM6CU_PGF_VAR=1 timeout 300 python -m pytest tests/test_bt_timeloop_gpu.py tests/test_pressure_force.py -m gpu -x -q 2>&1 | tail -2 ) > gpurun_out/r02_vars_parity.log; cat gpurun_out/r02_vars_parity.log
