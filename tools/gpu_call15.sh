python tools/mgpu_phases.py 2>&1 >/dev/null | grep -E "PF_PHASES rank 0 (it|summary)" | cut -c1-230 > gpurun_out/r02n_phases_n1.txt
python tools/td_iter_profile.py bgm_w260 > gpurun_out/r02n_td_bgm.out 2> gpurun_out/r02n_td_bgm.txt
python tools/td_iter_profile.py sv0_w220 2>&1 >/dev/null | grep -E "^sv0" > gpurun_out/r02n_td_sv0.txt
cat gpurun_out/r02n_phases_n1.txt; grep -E "^bgm|iteration (1|9|13|16|20):|Error|error" gpurun_out/r02n_td_bgm.txt | cut -c1-220; tail -n 3 gpurun_out/r02n_td_bgm.out; grep -c moved gpurun_out/r02n_td_bgm.txt; cat gpurun_out/r02n_td_sv0.txt
