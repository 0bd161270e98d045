for v in product noval novalnb nb26 unroll product; do
  if [ "$v" = product ]; then unset PF_ROUTER_LIB; else export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_$v.so"; fi
  echo "== $v"; python tools/mgpu_phases.py 2>&1 >/dev/null | grep PF_PHASES | awk '{printf "%s ", $0; if (NR>1) printf "\n"}' | grep -E "it  ?[0-9]|summary" | sed 's/PF_PHASES rank 0 //'
done > gpurun_out/r02e_variants_phases.txt 2>&1
unset PF_ROUTER_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file gpurun_out/r02e_gen_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python -m pytest tests/test_gpu_fullsize.py -q -s > gpurun_out/r02e_fullsize.txt 2>&1; echo "rc=$?" >> gpurun_out/r02e_fullsize.txt
cat gpurun_out/r02e_variants_phases.txt; grep -v "^==" gpurun_out/r02e_gen_launches.csv | cut -d, -f5,15 | tail -n 9; tail -n 6 gpurun_out/r02e_fullsize.txt
