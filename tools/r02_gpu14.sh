#!/bin/bash
# round 2, call 14: CorAdCalc at 4 CTAs/SM by default: its parity tests, the step tests, then the bench line of the round
mkdir -p gpurun_out
python -m pytest tests/test_coradcalc.py tests/test_reference_golden.py tests/test_step_dyn.py tests/test_benchmark_size_gpu.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_corad4_tests.log; cat gpurun_out/r02_corad4_tests.log
( timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err )
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "launches", d["gpu_launches"])
e = d["e2e"]; print("e2e", e["value"], "full", (e.get("full_cycle") or {}).get("value"), "host_state", (e.get("host_state_every_step") or {}).get("value"))
print({k: (round(v["ms_per_step"], 2), round(v.get("frac_of_peak", 0), 3)) for k, v in d["in_step"].items()})
print(d["state_checksum_after_steps"]["fields"]["h"]["bitcount"], d["clocks"])
PY
