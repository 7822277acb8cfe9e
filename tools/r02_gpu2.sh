#!/bin/bash
# round 2, GPU call 2: RECONSTRUCT_FOR_PRESSURE parity + timing
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_pressure_force.py -m gpu -x -q > gpurun_out/r02_pgf_recon.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pgf_recon.log )
for rs in 0 1 2; do ( MOM6CU_PGF_RECON=$rs timeout 300 python tools/prof_stage.py pgf 1440 1080 75 3 > gpurun_out/r02_pgf_time_$rs.log 2>&1 ); done
tail -12 gpurun_out/r02_pgf_recon.log; tail -3 gpurun_out/r02_pgf_time_*.log
