python tools/mgpu_phases.py 2>&1 >/dev/null | grep -E "PF_PHASES rank 0 (it|summary)" | cut -c1-230 > gpurun_out/r02m_phases_n1.txt
python tools/td_iter_profile.py bgm_w260 2>&1 >/dev/null | grep -E "iteration [0-9]+:|^bgm|moved|PF_PHASES rank 0 it" | cut -c1-200 > gpurun_out/r02m_td_bgm.txt
python tools/td_iter_profile.py sv0_w220 2>&1 >/dev/null | grep -E "^sv0" > gpurun_out/r02m_td_sv0.txt
cat gpurun_out/r02m_phases_n1.txt; grep -E "^bgm|iteration (1|9|13|16|20):" gpurun_out/r02m_td_bgm.txt; grep -c moved gpurun_out/r02m_td_bgm.txt; grep -E "it (9|13|16|20):" gpurun_out/r02m_td_bgm.txt; cat gpurun_out/r02m_td_sv0.txt
