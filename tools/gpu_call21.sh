export PF_LAUNCH_LOG=1
python tools/td_iter_profile.py bgm_w260 big_slots=64 > /dev/null 2> gpurun_out/r02p_bgm_64.txt
python tools/td_iter_profile.py bgm_w260 > /dev/null 2> gpurun_out/r02p_bgm_296.txt
grep -E "^bgm" gpurun_out/r02p_bgm_64.txt gpurun_out/r02p_bgm_296.txt | cut -c1-200
