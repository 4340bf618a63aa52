# round 2: parity at the configuration sizes: live reference harness (C1 170 k, DB 2 M), production vs IEEE kernel at DB(20 M), waiver-free tolerances
nvidia-smi -L
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "live_reference or at_20m or vs_oracle_steps or stages_vs_reference") > gpurun_out/pytest_live.log 2>&1; grep -v "^$" gpurun_out/pytest_live.log | cut -c1-400 | tail -40
