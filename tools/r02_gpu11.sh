#!/bin/bash
# round 2, GPU call 11 (gpurun --gpus 8): the bench at N = 8 on the full 1440x1080x75 problem, as the driver launches it (4x2 tiles)
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu \
    > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err )
tail -c 300 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: v["bitcount"] for k, v in d["state_checksum_after_steps"]["fields"].items()})
    print({k: round(v["ms_per_step"], 2) for k, v in d["in_step"].items()})
except Exception as e:
    print("unreadable", e)
PY
