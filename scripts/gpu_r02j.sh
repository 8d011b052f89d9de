#!/bin/bash
# Round-2j (2 GPUs): data-parallel backward captured into a CUDA graph (with its all-reduces) vs eager.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 "$@"; }
for tag in graph eager graph2; do
  if [ $tag = eager ]; then export FVIT_DDP_GRAPH=0; else unset FVIT_DDP_GRAPH; fi
  t bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02j_bench_n2_$tag.json 2> gpurun_out/r02j_bench_n2_$tag.err
  echo "n2 $tag exit $?"; grep -v "^NCCL" gpurun_out/r02j_bench_n2_$tag.json | cut -c1-200; grep -i "warn\|error\|capture" gpurun_out/r02j_bench_n2_$tag.err | head -5 | cut -c1-300
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02j_bench_n2_$tag.json') if l.startswith('{')][-1]); print('  ', d['value'], d['ms_per_step'], 'grad_sync', d['grad_sync_check'], 'launches/step', d['launches_per_step'])"
done
python bench.py --steps 10 --warmup 3 --no-e2e --no-also > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02j_bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q -k "21k" -p no:cacheprovider 2>&1 | tail -3
