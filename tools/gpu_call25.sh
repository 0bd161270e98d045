export PF_LAUNCH_LOG=1
python tools/td_iter_profile.py bgm_w260 > /dev/null 2> gpurun_out/r02t_bgm.txt
python tools/td_iter_profile.py bgm_w260 big_slots=64 > /dev/null 2> gpurun_out/r02t_bgm_64.txt
python tools/td_iter_profile.py sv0_w220 > /dev/null 2> gpurun_out/r02t_sv0.txt
export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_diag.so"
python tools/td_iter_profile.py bgm_w260 big_slots=64 > /dev/null 2> gpurun_out/r02t_bgm_diag.txt
grep -E "^(bgm|sv0)" gpurun_out/r02t_*.txt | cut -c1-220
