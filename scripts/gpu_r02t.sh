#!/bin/bash
# Round-2t: long-window attention backward with eight softmax warps + bias prefetch: kernel tests, tiny_21k training parity,
# micro-benchmark on the 21k-384 shape and on the two-tile shapes (any-res S = 148, 21k-224 S = 196) against the two-tile kernel.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider \
  -k "long or 21k" > gpurun_out/r02t_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error|error" gpurun_out/r02t_pytest.log | tail -25
( timeout 120 python scripts/gpu_attn_long_micro.py; timeout 120 python scripts/gpu_attn_long_micro.py 148 8 32 960; \
  timeout 120 python scripts/gpu_attn_long_micro.py 196 16 49 128 ) > gpurun_out/r02t_attn_long_micro.txt 2>&1
cut -c1-220 gpurun_out/r02t_attn_long_micro.txt
