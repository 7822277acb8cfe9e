#!/bin/bash
# round 2, call 15: ncu --set full of the other stage kernels AT FULL SIZE inside a real step (the round-1 captures were at 720x540)
mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'hor_visc_kernel|corad_kernel|pgf_recon_kernel|pgf_ts_edges|vv_coef_kernel|vv_solve_kernel|bt_col_kernel|bt_layer_accel' \
    -c 26 -f -o gpurun_out/r02_stage_kernels_fullsize \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-stages --no-thermo > gpurun_out/r02_stage_kernels_ncu.log 2>&1 )
tail -3 gpurun_out/r02_stage_kernels_ncu.log
ls -la gpurun_out/r02_stage_kernels_fullsize.ncu-rep
