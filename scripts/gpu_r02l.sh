#!/bin/bash
# Round-2l (2 GPUs): data-parallel backward as CUDA-graph segments; clean exit check.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 "$@"; }
for tag in seg eager; do
  if [ $tag = eager ]; then export FVIT_CUDA_GRAPH=0; else unset FVIT_CUDA_GRAPH; fi
  t0=$(date +%s)
  t bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02l_bench_n2_$tag.json 2> gpurun_out/r02l_bench_n2_$tag.err
  echo "n2 $tag exit $? after $(( $(date +%s) - t0 ))s"; grep -i "warn\|capture\|error" gpurun_out/r02l_bench_n2_$tag.err | head -4 | cut -c1-300
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02l_bench_n2_$tag.json') if l.startswith('{')][-1]); print('  ', d['value'], d['ms_per_step'], 'grad_sync', d['grad_sync_check'], 'launches/step', d['launches_per_step'])"
done
unset FVIT_CUDA_GRAPH
python bench.py --steps 10 --warmup 3 --no-e2e --no-also > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02l_bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
