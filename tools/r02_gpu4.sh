#!/bin/bash
# round 2, GPU call 4: the resident-cycle e2e leg at quarter size; ncu capture of the PressureForce reconstruction kernel
mkdir -p gpurun_out
( timeout 500 python bench.py --size 720,540 --steps 3 --warmup 1 --no-cpu --no-stages --no-thermo > gpurun_out/r02_cycle_small.json 2> gpurun_out/r02_cycle_small.err )
tail -c 600 gpurun_out/r02_cycle_small.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_cycle_small.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", json.dumps(d["e2e"])[:1500])
except Exception as e:
    print("unreadable", e)
PY
( MOM6CU_PGF_RECON=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgf_recon -c 1 -f -o gpurun_out/r02_pgf_recon \
    python tools/prof_stage.py pgf 720 540 75 1 > gpurun_out/r02_pgf_recon_ncu.log 2>&1 )
tail -3 gpurun_out/r02_pgf_recon_ncu.log
