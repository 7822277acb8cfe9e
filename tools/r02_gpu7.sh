#!/bin/bash
# round 2, GPU call 7: DRAM traffic of every continuity kernel at full size (ncu --set full), the launch list of the bench command, btstep timing
mkdir -p gpurun_out
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:cont_ -c 4 -f -o gpurun_out/r02_cont_all \
    python tools/prof_cont.py 1440 1080 75 1 > gpurun_out/r02_cont_all_ncu.log 2>&1 )
tail -2 gpurun_out/r02_cont_all_ncu.log
( timeout 300 python tools/prof_stage.py btstep 1440 1080 75 3 2>&1 | tail -2 ) > gpurun_out/r02_btstep_time.log; cat gpurun_out/r02_btstep_time.log
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_bench_1440x1080.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-stages --no-thermo > gpurun_out/r02_launches_bench.log 2>&1 )
tail -c 300 gpurun_out/r02_launches_bench.log; wc -l gpurun_out/r02_launches_bench_1440x1080.csv
