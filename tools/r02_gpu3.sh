#!/bin/bash
# round 2, GPU call 3: the whole GPU suite after the device pow / PGF reconstruction / bench changes, then the bench line
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gpu_suite.log )
( timeout 600 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err )
tail -6 gpurun_out/r02_gpu_suite.log; tail -c 1500 gpurun_out/r02_bench1.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench1.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "e2e", "state_checksum_after_steps")})
    print({k: (round(v["ms_per_step"], 2), round(v.get("frac_of_peak", 0), 3)) for k, v in d["in_step"].items()})
except Exception as e:
    print("bench line unreadable", e)
PY
