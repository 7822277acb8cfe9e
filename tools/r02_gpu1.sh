#!/bin/bash
# round 2, GPU call 1: benchmark-size parity tests, baseline bench, full-size ncu capture of cont_flux
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_benchmark_size_gpu.py -m gpu -x -q > gpurun_out/r02_benchsize.log 2>&1; echo "rc=$?" >> gpurun_out/r02_benchsize.log )
( timeout 400 python bench.py > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err )
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:cont_flux -c 2 -f -o gpurun_out/r02_cont_full \
    python tools/prof_cont.py 1440 1080 75 1 > gpurun_out/r02_cont_ncu.log 2>&1 )
tail -15 gpurun_out/r02_benchsize.log; cat gpurun_out/r02_bench0.json; nproc; free -g | head -2
