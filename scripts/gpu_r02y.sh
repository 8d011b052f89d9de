#!/bin/bash
# Round-2y: full GPU suite after the qk_scale plumbing (tiny_qk golden: eval logits, per-module activations, training step).
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r02y_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02y_pytest.log | tail -24
