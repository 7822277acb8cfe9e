#!/bin/bash
# round 2, GPU call 6: the whole step at the HEADLINE grid (1440x1080x75, the bench's land fraction and PLM pressure reconstruction) against the
# oracle, bit for bit; and the benchmark-size cases with the pressure reconstruction
mkdir -p gpurun_out
( MOM6CU_TEST_FULL_SIZE=1 timeout 2400 python -m pytest tests/test_benchmark_size_gpu.py -m gpu -q -k "headline or reconstruction" --durations=5 \
    > gpurun_out/r02_headline_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r02_headline_parity.log )
tail -12 gpurun_out/r02_headline_parity.log
