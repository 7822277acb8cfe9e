#!/bin/bash
# First GPU call of the next round (gpurun --timeout 900 -- 'bash tools/round2_gpu_first.sh'): what the round-1 GPU budget could not cover.
#  1. the GPU tests written after the budget was spent (sorted last in the suite): the FGNV / stored-slope / MEKE selection of
#     thickness_diffuse, the Eady / MEKE terms of tracer_hordiff, mom6cu_do_group_pass and the chained callers;
#  2. full-size timings of thickness_diffuse and tracer_hordiff (tools/time_callers.py) and of mixedlayer_restrat (tools/time_mle.py);
#  3. a --set full capture of their kernels at 720x540x75 for profiles/ (read here with tools/ncu_summary.py).
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_zz2_thickness_diffuse_ext_gpu.py tests/test_zz3_tracer_hordiff_ext_gpu.py tests/test_zz1_mle_ext_gpu.py tests/test_zz4_callers_chain_gpu.py tests/test_zz5_callers_full_size_gpu.py -m gpu -q \
    > gpurun_out/r02_gpu_zz.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gpu_zz.log )
( timeout 120 python tools/time_callers.py > gpurun_out/r02_time_callers.json 2> gpurun_out/r02_time_callers.err )
( timeout 90 python tools/time_mle.py > gpurun_out/r02_time_mle.json 2> gpurun_out/r02_time_mle.err )
( timeout 150 ncu --set full --clock-control none --import-source on -k regex:'td_|hd_' -c 12 -f -o gpurun_out/r02_callers_ncu \
    python tools/time_callers.py 720,540,75 > gpurun_out/r02_callers_ncu.log 2>&1 )
tail -5 gpurun_out/r02_gpu_zz.log; cat gpurun_out/r02_time_callers.json gpurun_out/r02_time_mle.json
