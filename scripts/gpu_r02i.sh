#!/bin/bash
# Round-2i evidence: ncu launch list (time + DRAM bytes) of one fv4 training + optimizer step, ncu --set full captures of
# the round-2 kernels (CTA-pair GEMM, fused HAT attention, key-loop attention fwd/bwd, attention bwd, BN / LN / affine).
mkdir -p gpurun_out
export FVIT_CUDA_GRAPH=0
t0=$(date +%s)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02i_launches_fv4_step.csv python scripts/gpu_train_step_profile.py \
    > gpurun_out/r02i_ncu_launches.log 2>&1
echo "ncu launch list exit $? after $(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/r02i_ncu_launches.log | cut -c1-200; wc -l gpurun_out/r02i_launches_fv4_step.csv
full() { # name kernel-regex skip count cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  local t1=$(date +%s)
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt \
      -o gpurun_out/r02i_$name "$@" > gpurun_out/r02i_ncu_$name.log 2>&1
  echo "ncu $name exit $? after $(( $(date +%s) - t1 ))s"; tail -1 gpurun_out/r02i_ncu_$name.log | cut -c1-160
}
full gemm_train 'gemm_tcgen05' 150 6 python scripts/gpu_train_step_profile.py
full attn_train 'attn_(bwd_)?tc_kernel' 20 4 python scripts/gpu_train_step_profile.py
full glue_train 'bn_bwd|ln_bwd_dx|ln_fwd|affine_rows' 30 8 python scripts/gpu_train_step_profile.py
cat > /tmp/fwd_prof.py <<'PY'
import sys, torch
sys.path.insert(0, '/root/repo')
import fastervit_b200 as F
name, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
kw = dict(resolution=[H, W], window_size=[7, 7, 12, 6], ct_size=2, dim=64) if 'any_res' in name else {}
m = F.create_model(name, **kw).cuda().eval()
x = torch.randn(B, 3, H, W, device='cuda')
with torch.no_grad():
    for _ in range(2): m(x)
    torch.cuda.synchronize(); torch.cuda.profiler.start(); m(x); torch.cuda.synchronize(); torch.cuda.profiler.stop()
PY
full hat_fv4_fwd 'hat_attn_kernel' 2 3 python /tmp/fwd_prof.py faster_vit_4_224 128 224 224
full loop_ar0_fwd 'attn_loop_kernel' 1 2 python /tmp/fwd_prof.py faster_vit_0_any_res 32 576 960
ls -la gpurun_out/*.ncu-rep 2>/dev/null | cut -c30-120
