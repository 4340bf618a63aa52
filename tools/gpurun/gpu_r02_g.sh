# round 2: gauges against the reference's gauge dumps + the whole GPU suite
nvidia-smi -L
(time timeout 600 python -m pytest tests/test_gauges.py -m gpu -x -q -s) > gpurun_out/pytest_gauges.log 2>&1; tail -30 gpurun_out/pytest_gauges.log | cut -c1-500
