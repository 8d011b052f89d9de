#!/bin/bash
# Round-2c: CTA-pair GEMM (templated), loop attention bwd, fused HAT kernel, CUDA graphs.
mkdir -p gpurun_out
run() { # name, timeout, args...
  local name=$1; local to=$2; shift 2
  local t0=$(date +%s)
  timeout $to "$@" > gpurun_out/r02c_$name.log 2>&1
  echo "$name exit $? after $(( $(date +%s) - t0 ))s"
  grep -E "passed|failed|FAILED|worst activation|median rel" gpurun_out/r02c_$name.log | cut -c1-230 | tail -${TAILN:-25}
}
export FVIT_CUDA_GRAPH=0
run gemm 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider
run attn 600 python -m pytest tests/test_attn_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider
timeout 300 python scripts/gpu_gemm_pair_micro.py > gpurun_out/r02c_pair_micro.log 2>&1
echo "micro exit $?"; cat gpurun_out/r02c_pair_micro.log | cut -c1-330
TAILN=45 run model_nograph 900 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py tests/test_train_ops_gpu.py -m gpu -q --maxfail=30 -p no:cacheprovider -s
unset FVIT_CUDA_GRAPH
TAILN=45 run model_graph 900 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q --maxfail=30 -p no:cacheprovider -s
