#!/bin/sh
# multi-GPU evidence in one call:  gpurun --gpus N -- 'sh tools/gpu_n.sh N tag [big]'
N="$1"; tag="$2"; big="$3"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR tools/mgpu_phases.py > gpurun_out/${tag}_phases_n$N.out 2> gpurun_out/${tag}_phases_n$N.txt; echo "phases rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
grep -a PF_PHASES gpurun_out/${tag}_phases_n$N.txt | sort -s -k3,3n | tail -n 80; tail -n 3 gpurun_out/${tag}_bench_n$N.err; cut -c1-700 gpurun_out/${tag}_bench_n$N.json
if [ -n "$big" ]; then
  timeout 900 $TR tools/mgpu_phases.py 800 800000 > gpurun_out/${tag}_phases800_n$N.out 2> gpurun_out/${tag}_phases800_n$N.txt; echo "phases800 rc=$?"
  grep -a "PF_PHASES rank [0-9]* summary" gpurun_out/${tag}_phases800_n$N.txt
fi
