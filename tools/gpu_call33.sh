# N ranks: the bench line incl. e2e with the cached transport (exchange regions + peer mappings kept across routers), then the phases
N="$1"; tag="$2"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
mkdir -p gpurun_out
timeout 150 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_bench_n$N.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['route']['iterations'], d['route']['wirelength'][:2], d['e2e']['value'], d['e2e']['s_per_step'], d['e2e']['phases_s'])"; grep -av "destroy_process_group\|^$" gpurun_out/${tag}_bench_n$N.err | tail -n 4
timeout 100 $TR tools/mgpu_phases.py > gpurun_out/${tag}_phases_n$N.out 2> gpurun_out/${tag}_phases_n$N.txt; echo "phases rc=$?"
grep -a "PF_PHASES rank [0-9]* summary" gpurun_out/${tag}_phases_n$N.txt | cut -c1-330
