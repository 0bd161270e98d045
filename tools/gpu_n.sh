#!/bin/sh
# multi-GPU evidence in one call:  gpurun --gpus N -- 'sh tools/gpu_n.sh N tag'
N="$1"; tag="$2"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR tools/mgpu_phases.py > gpurun_out/${tag}_phases_n$N.out 2> gpurun_out/${tag}_phases_n$N.txt; echo "phases rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
grep PF_PHASES gpurun_out/${tag}_phases_n$N.txt | tail -n 40; tail -n 3 gpurun_out/${tag}_bench_n$N.err; cat gpurun_out/${tag}_bench_n$N.json | cut -c1-1500
