#!/bin/bash
# Round-2v (2 GPUs): data-parallel fv4 training step after the round's last changes (wait_side markers in the launch
# list, weight gradients on the main branch under the all-reduce) -- N=1 on the same box, N=2 default, N=2 with the
# weight-gradient side branch forced on (FVIT_WGRAD_SIDE_DDP=1).
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
show() { python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02v_bench_$1.json').read().strip().splitlines()[-1])
    print('  $1', d['value'], 'img/s', d['ms_per_step'], 'ms', d.get('data_parallel'), 'grad_sync', d.get('grad_sync_check'))
except Exception as e: print('  $1 no line', e)
PY
}
timeout 300 python bench.py --workload fv4_train --steps 10 --warmup 3 --no-also --no-e2e > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; show n1
b2() { # tag [env...]
  local tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload fv4_train --steps 10 --warmup 3 --no-also --no-e2e > gpurun_out/r02v_bench_$tag.json 2> gpurun_out/r02v_bench_$tag.err
  show $tag
}
b2 n2_default FVIT_WGRAD_SIDE_DDP=0
b2 n2_wgrad_side FVIT_WGRAD_SIDE_DDP=1
tail -3 gpurun_out/r02v_bench_n2_default.err | cut -c1-300
