#!/bin/bash
# Round-2b: fixed parity tests + CTA-pair GEMM tests and micro-benchmark.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --maxfail=40 -p no:cacheprovider -k "pair" > gpurun_out/r02b_pytest_pair.log 2>&1
echo "pair pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02b_pytest_pair.log | cut -c1-220 | tail -40
timeout 300 python -m pytest tests/test_attn_gpu.py -m gpu -q --maxfail=40 -p no:cacheprovider -k "loop" -s > gpurun_out/r02b_pytest_loop.log 2>&1
echo "loop attn pytest exit $?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r02b_pytest_loop.log | cut -c1-220 | tail -30
timeout 300 python scripts/gpu_gemm_pair_micro.py > gpurun_out/r02b_pair_micro.log 2>&1
echo "micro exit $?"; cat gpurun_out/r02b_pair_micro.log | cut -c1-330
t0=$(date +%s)
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py tests/test_train_ops_gpu.py -m gpu -q --maxfail=30 -p no:cacheprovider -s > gpurun_out/r02b_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error|worst activation|median rel" gpurun_out/r02b_pytest.log | cut -c1-260 | tail -60
