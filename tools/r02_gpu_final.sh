#!/bin/bash
# round 2, final GPU call: the whole GPU suite, smoke(), the bench line as the driver runs it, CorAdCalc 3 vs 4 CTAs/SM
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_suite_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gpu_suite_final.log )
tail -5 gpurun_out/r02_gpu_suite_final.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02_smoke.log ); tail -2 gpurun_out/r02_smoke.log
for m in 3 4; do ( MOM6CU_CORAD_MINB=$m timeout 200 python tools/prof_stage.py corad 1440 1080 75 3 2>&1 | tail -1 | sed "s/^/MINB=$m /" ) >> gpurun_out/r02_corad_minb.log; done; cat gpurun_out/r02_corad_minb.log
( timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err )
tail -c 400 gpurun_out/r02_bench_final.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
    print("value", d["value"], d["ms_per_step"], "launches", d["gpu_launches"])
    e = d["e2e"]; print("e2e", e["value"], "full", (e.get("full_cycle") or {}).get("value"), "host_state", (e.get("host_state_every_step") or {}).get("value"))
    print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    print({k: (round(v["ms_per_step"], 2), round(v.get("frac_of_peak", 0), 3)) for k, v in d["in_step"].items()})
    print(d["btstep_microbench"]["frac_of_peak"], d["state_checksum_after_steps"]["fields"]["h"]["bitcount"])
except Exception as ex:
    print("bench line unreadable", ex)
PY
( timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null ); tail -c 600 gpurun_out/r02_bench_reference_arm.json
