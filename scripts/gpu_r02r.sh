#!/bin/bash
# Round-2r: long-window attention backward (S > 256: fvit_attn_loop_bwd_long) -- kernel tests, tiny_21k training parity,
# then the two new side measurements of the bench (fv0 batch-8 latency, faster_vit_4_21k_384 training step).
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider \
  -k "long or loop or 21k" > gpurun_out/r02r_pytest.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - t0 ))s"; grep -E "passed|failed|FAILED|Error|error" gpurun_out/r02r_pytest.log | tail -25
timeout 400 python - > gpurun_out/r02r_side.log 2>&1 <<'PY'
import json, torch, bench
pk = bench.peaks()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
for wl in ("fv4_21k_384_train", "fv0_fwd_b8"):
    try:
        print(wl, json.dumps(bench.quick_measure(wl, dev, pk)), flush=True)
    except Exception as e:
        import traceback; traceback.print_exc()
PY
echo "side exit $?"; cut -c1-900 gpurun_out/r02r_side.log | tail -12
