# where create_router spends its time on N ranks (PF_COMM_DEBUG: verbose library lines + the python-side steps)
N="$1"; tag="$2"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
mkdir -p gpurun_out
PF_COMM_DEBUG=1 timeout 150 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_commdebug_n$N.out 2> gpurun_out/${tag}_commdebug_n$N.err; echo "bench rc=$?"
grep -a "create_router\|connect:\|pf_router: create\|stripes" gpurun_out/${tag}_commdebug_n$N.out gpurun_out/${tag}_commdebug_n$N.err | tail -n 24
