#!/bin/bash
# Launch bench.py under torchrun exactly like the round-end driver does; logs go to gpurun_out/.
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/ddp_gpus.txt 2>&1
python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "torchrun exit code $?" >> gpurun_out/bench_n$N.err
echo "--- stdout"; cut -c1-400 gpurun_out/bench_n$N.json
echo "--- stderr tail"; tail -25 gpurun_out/bench_n$N.err | cut -c1-300
echo "--- gpus"; cat gpurun_out/ddp_gpus.txt
