#!/bin/bash
# Round-2h (2 GPUs): data-parallel timeline with / without the all-reduce SM cap, 2-GPU bench lines; ln_bwd prefetch A/B.
mkdir -p gpurun_out
export FVIT_BENCH_CPU_BUDGET_S=2
t() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
t scripts/gpu_ddp_timeline.py gpurun_out/r02h_ddp_timeline.json > gpurun_out/r02h_timeline.log 2>&1
echo "timeline exit $?"; tail -60 gpurun_out/r02h_timeline.log | cut -c1-200
FVIT_NCCL_CTAS=0 NCCL_MAX_CTAS=32 t bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02h_bench_n2_nocap.json 2> gpurun_out/r02h_bench_n2_nocap.err
echo "n2 nocap exit $?"; cut -c1-330 gpurun_out/r02h_bench_n2_nocap.json
t bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02h_bench_n2_cap.json 2> gpurun_out/r02h_bench_n2_cap.err
echo "n2 cap exit $?"; cut -c1-330 gpurun_out/r02h_bench_n2_cap.json
python bench.py --steps 10 --warmup 3 --no-e2e --no-also > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
echo "n1 exit $?"; cut -c1-330 gpurun_out/r02h_bench_n1.json
FVIT_LN_PREFETCH=1 python bench.py --steps 10 --warmup 3 --no-e2e --no-also > gpurun_out/r02h_bench_n1_lnp.json 2> gpurun_out/r02h_bench_n1_lnp.err
echo "n1 ln prefetch exit $?"; python -c "
import json
for f in ('n1','n1_lnp'):
    d=json.loads(open(f'gpurun_out/r02h_bench_{f}.json').read().strip().splitlines()[-1]); pk=d['per_kernel']
    print(f, d['value'], d['ms_per_step'], pk['fvit_ln_bwd']['ms'])
"
