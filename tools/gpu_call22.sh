export PF_LAUNCH_LOG=1
python tools/td_iter_profile.py bgm_w260 > /dev/null 2> gpurun_out/r02q_bgm.txt
python tools/td_iter_profile.py bgm_w260 big_slots=64 > /dev/null 2> gpurun_out/r02q_bgm_64.txt
python tools/td_iter_profile.py bgm_w260 lazy_seed_min=-1 > /dev/null 2> gpurun_out/r02q_bgm_eager.txt
python tools/td_iter_profile.py sv0_w220 > /dev/null 2> gpurun_out/r02q_sv0.txt
grep -E "^(bgm|sv0)" gpurun_out/r02q_*.txt | cut -c1-220
