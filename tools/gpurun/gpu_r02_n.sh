# round 2: whole GPU suite after mDBC / gauges / drop-in
nvidia-smi -L | head -1
(time timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
