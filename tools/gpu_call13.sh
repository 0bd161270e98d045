for v in product farinl product; do
  if [ "$v" = product ]; then unset PF_ROUTER_LIB; else export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_$v.so"; fi
  echo "== $v"; python tools/mgpu_phases.py 2>&1 >/dev/null | grep -E "PF_PHASES rank 0 (it  1|summary)" | cut -c1-230
done > gpurun_out/r02l_variants_phases.txt 2>&1
unset PF_ROUTER_LIB
python tools/td_iter_profile.py bgm_w260 2>&1 >/dev/null | grep -E "iteration [0-9]+:|^bgm|moved" | cut -c1-200 > gpurun_out/r02l_td_bgm.txt
python tools/td_iter_profile.py sv0_w220 2>&1 >/dev/null | grep -E "^sv0" > gpurun_out/r02l_td_sv0.txt
cat gpurun_out/r02l_variants_phases.txt; grep -E "^bgm|iteration (1|9|13|16|20):" gpurun_out/r02l_td_bgm.txt; grep -c moved gpurun_out/r02l_td_bgm.txt; cat gpurun_out/r02l_td_sv0.txt
