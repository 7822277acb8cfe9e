#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus 8): layout-invariance tests on 2, 4 and 8 GPUs (2x1, 1x2, 2x2 with corner exchanges, 4x2, 2x4),
# then the bench's state checksums at N = 1, 2, 4, 8 on ONE global problem (720x540x75 to keep the call short; the driver's own scaling
# run does the same at 1440x1080x75)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_multi_gpus.txt
( timeout 900 python -m pytest tests/test_step_multigpu.py tests/test_bt_multigpu.py tests/test_diag.py tests/test_zz4_callers_chain_gpu.py -m gpu -q -rs \
    > gpurun_out/r02_multigpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_multigpu_tests.log )
tail -8 gpurun_out/r02_multigpu_tests.log
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    ( timeout 300 python bench.py --gpus 1 --steps 3 --warmup 1 --size 720,540 --no-cpu --no-stages --no-thermo --no-e2e > gpurun_out/r02_scale_small_$n.json 2> gpurun_out/r02_scale_small_$n.err )
  else
    ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 1 \
        --size 720,540 --no-cpu --no-stages --no-thermo --no-e2e > gpurun_out/r02_scale_small_$n.json 2> gpurun_out/r02_scale_small_$n.err )
  fi
  tail -c 400 gpurun_out/r02_scale_small_$n.err
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r02_scale_small_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"] / 1e6, 1), "M cell-updates/s", d["ms_per_step"], {k: v["bitcount"] for k, v in d["state_checksum_after_steps"]["fields"].items()})
    except Exception as e:
        print(n, "unreadable", e)
PY
