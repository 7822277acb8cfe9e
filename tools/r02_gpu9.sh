#!/bin/bash
# round 2, GPU call 9 (gpurun --gpus 2): the bench at N = 2 on the full 1440x1080x75 problem, exactly as the driver launches it
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err )
tail -c 500 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: v["bitcount"] for k, v in d["state_checksum_after_steps"]["fields"].items()})
    print({k: round(v["ms_per_step"], 2) for k, v in d["in_step"].items()})
except Exception as e:
    print("unreadable", e)
PY
free -g | head -2
