# new defaults (in-flight divisor 32 on tight fabrics, ripple floor 1024) against the old ones: whole GPU suite, run-to-run spread, the
# two BASELINE-scale timing-driven stand-ins, and the cfg 4 bench line (must be unchanged: its channels are under 40 % full)
python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.txt
tail -n 5 gpurun_out/r02z_pytest.txt
python tools/parity_repeat.py 6 heq_w70,het_w70,mix_w70,toy_w64,hub_w90 "" "inflight_div=16 ripple_max_nets=256" 2>&1 | tee gpurun_out/r02z_repeat.txt
python tools/parity_repeat.py 2 mid_w200 "" "inflight_div=16 ripple_max_nets=256" 2>&1 | tee -a gpurun_out/r02z_repeat.txt
python tools/td_profile.py sv0_w220 '{}' q 2>&1 | tee gpurun_out/r02z_td.txt
python tools/td_profile.py sv0_w220 '{"inflight_div":16,"ripple_max_nets":256}' q 2>&1 | tee -a gpurun_out/r02z_td.txt
python tools/td_profile.py bgm_w260 '{}' q 2>&1 | tee -a gpurun_out/r02z_td.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; cut -c1-400 gpurun_out/r02z_bench.json
PF_PHASES=1 python tools/mgpu_phases.py 800 800000 2> gpurun_out/r02z_phases800_n1.txt; grep -a summary gpurun_out/r02z_phases800_n1.txt
