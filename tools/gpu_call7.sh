for v in product refunroll product; do
  if [ "$v" = product ]; then unset PF_ROUTER_LIB; else export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_$v.so"; fi
  echo "== $v"; python tools/mgpu_phases.py 2>&1 >/dev/null | grep PF_PHASES | sed 's/PF_PHASES rank 0 //'
done > gpurun_out/r02f_variants_phases.txt 2>&1
unset PF_ROUTER_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file gpurun_out/r02f_gen_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_route_kernel -c 1 -o gpurun_out/r02f_route -f python tools/prof_run.py 400 200000 1 > gpurun_out/r02f_prof_run.log 2>&1
grep -E "^==|it  1|summary" gpurun_out/r02f_variants_phases.txt | cut -c1-250; grep -v "^==" gpurun_out/r02f_gen_launches.csv | tail -n 8 | cut -d'"' -f10,30; tail -n 2 gpurun_out/r02f_prof_run.log
