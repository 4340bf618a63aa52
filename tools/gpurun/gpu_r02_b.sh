# round 2: parity suite after the refbinary-tail switch + occupancy / ring-depth alternatives of variant 10 at DB(20 M)
nvidia-smi -L
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
(time timeout 400 python tools/forcebench.py 20000000 8 10,11,12,13,14,15,10,11,12) > gpurun_out/forcebench_v10alts.log 2>&1; tail -12 gpurun_out/forcebench_v10alts.log
