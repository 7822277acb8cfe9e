#!/bin/bash
# round 2, call 13: vertvisc_limit_vel on the device (oracle parity + the reference digests) and the suites it touches
mkdir -p gpurun_out
python -m pytest tests/test_vertvisc.py tests/test_reference_golden.py tests/test_step_dyn.py tests/test_golden_digests.py tests/test_abi.py tests/test_benchmark_size_gpu.py -q -m gpu --durations=5 2>&1 | tail -25 > gpurun_out/r02_limit_vel_gpu.log
cat gpurun_out/r02_limit_vel_gpu.log
