export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_diag.so"
python tools/td_iter_profile.py bgm_w260 big_slots=64 > /dev/null 2> gpurun_out/r02s_bgm_diag.txt
grep -E "^(bgm|sv0)" gpurun_out/r02s_*.txt | cut -c1-220
