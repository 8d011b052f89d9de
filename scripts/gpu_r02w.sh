#!/bin/bash
# Round-2w evidence refresh for the final tree: ncu launch list (time + DRAM bytes) of one fv4 training + optimizer step
# (weight gradients on the side branch are serialised by ncu: shares, not overlap), ncu --set full of the eight-softmax-warp
# long-window attention backward and of two row-per-thread-epilogue GEMM launches.
mkdir -p gpurun_out
export FVIT_CUDA_GRAPH=0
t0=$(date +%s)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02w_launches_fv4_step.csv python scripts/gpu_train_step_profile.py \
    > gpurun_out/r02w_ncu_launches.log 2>&1
echo "ncu launch list exit $? after $(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/r02w_ncu_launches.log | cut -c1-200; wc -l gpurun_out/r02w_launches_fv4_step.csv
t1=$(date +%s)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_loop_bwd_long -s 2 -c 1 \
    -o gpurun_out/r02w_attn_long python scripts/gpu_attn_long_micro.py > gpurun_out/r02w_ncu_attn_long.log 2>&1
echo "ncu attn_long exit $? after $(( $(date +%s) - t1 ))s"
t2=$(date +%s)
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 150 -c 4 \
    -o gpurun_out/r02w_gemm_train python scripts/gpu_train_step_profile.py > gpurun_out/r02w_ncu_gemm.log 2>&1
echo "ncu gemm exit $? after $(( $(date +%s) - t2 ))s"
ls -la gpurun_out/*.ncu-rep 2>/dev/null | cut -c30-120
