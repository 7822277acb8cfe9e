#!/bin/bash
# round 2, call 12: the CUDA path against digests of the reference's own outputs (tests/test_reference_golden.py) + the btstep refusal
mkdir -p gpurun_out
python -m pytest tests/test_reference_golden.py -q -m gpu --durations=5 2>&1 | tail -25 > gpurun_out/r02_reference_golden_gpu.log
echo "rc=$?" >> gpurun_out/r02_reference_golden_gpu.log
cat gpurun_out/r02_reference_golden_gpu.log

