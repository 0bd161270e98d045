# parity after the round's last kernel changes: the whole GPU suite (no -x) and the run-to-run spread of the tight fixtures
python -m pytest tests -m gpu -q > gpurun_out/r02y_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y_pytest.txt
tail -n 5 gpurun_out/r02y_pytest.txt
python tools/parity_repeat.py 6 heq_w70 "" "inflight_div=32" "inflight_div=64" "polish=1" "ripple_max_nets=100000" "validate_commits=4" 2>&1 | tee gpurun_out/r02y_repeat.txt
python tools/parity_repeat.py 4 het_w70,mix_w70,toy_w64,hub_w90 "" "inflight_div=32" 2>&1 | tee -a gpurun_out/r02y_repeat.txt
