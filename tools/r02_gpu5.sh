#!/bin/bash
# round 2, GPU call 5: PGF parity after the code-size change + timing, continuity 2 vs 3 CTAs/SM, the bench line with the resident e2e legs
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_pressure_force.py -m gpu -x -q > gpurun_out/r02_pgf_recon2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pgf_recon2.log )
tail -3 gpurun_out/r02_pgf_recon2.log
( MOM6CU_PGF_RECON=1 timeout 300 python tools/prof_stage.py pgf 1440 1080 75 3 2>&1 | tail -2 ) > gpurun_out/r02_pgf_time_v2.log
cat gpurun_out/r02_pgf_time_v2.log
( timeout 300 python tools/prof_cont.py 1440 1080 75 3 2>&1 | tail -2 ) > gpurun_out/r02_cont_minb2.log
( MOM6CU_CONT_MINB=3 timeout 300 python tools/prof_cont.py 1440 1080 75 3 2>&1 | tail -2 ) > gpurun_out/r02_cont_minb3.log
cat gpurun_out/r02_cont_minb2.log gpurun_out/r02_cont_minb3.log
( timeout 900 python bench.py > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err )
tail -c 600 gpurun_out/r02_bench2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench2.json").read().strip().splitlines()[-1])
    print("value", d["value"], d["ms_per_step"])
    e = d["e2e"]; print("e2e", e["value"], e["h2d_bytes_per_step"], "full", (e.get("full_cycle") or {}).get("value"), "host_state", (e.get("host_state_every_step") or {}).get("value"))
    print({k: (round(v["ms_per_step"], 2), round(v.get("frac_of_peak", 0), 3)) for k, v in d["in_step"].items()})
except Exception as ex:
    print("bench line unreadable", ex)
PY
