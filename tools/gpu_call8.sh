# 1 GPU: tie-break A/B, the two BASELINE-scale timing-driven workloads (ours and the reference arm), default bench line
sh tools/ab_multi.sh 2 tienode > gpurun_out/r02g_ab_tienode.txt 2>&1
python bench.py --workload sv0 --steps 5 --warmup 2 > gpurun_out/r02g_bench_sv0.json 2> gpurun_out/r02g_bench_sv0.err
python bench.py --workload bgm --steps 5 --warmup 2 > gpurun_out/r02g_bench_bgm.json 2> gpurun_out/r02g_bench_bgm.err
python bench.py --impl reference --workload sv0 --steps 1 --warmup 0 > gpurun_out/r02g_ref_sv0.json 2> gpurun_out/r02g_ref_sv0.err
python bench.py --impl reference --workload bgm --steps 1 --warmup 0 > gpurun_out/r02g_ref_bgm.json 2> gpurun_out/r02g_ref_bgm.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
cat gpurun_out/r02g_ab_tienode.txt; for f in sv0 bgm; do cut -c1-900 gpurun_out/r02g_bench_$f.json; tail -n 2 gpurun_out/r02g_bench_$f.err; cut -c1-600 gpurun_out/r02g_ref_$f.json; done; cut -c1-200 gpurun_out/r02g_bench.json
