#!/bin/bash
# adapter example replays + owned-row heat kernel (1 GPU part of the run; the 2-GPU part is gpu_round2_v.sh)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_adapter_gpu.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/r2u_tests.log 2>&1
cat gpurun_out/r2u_tests.log
