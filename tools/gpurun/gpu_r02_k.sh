# round 2: the compiled drop-in on C1
nvidia-smi -L | head -1
(time timeout 900 python -m pytest tests/test_dropin.py -m gpu -x -q -s) > gpurun_out/pytest_dropin.log 2>&1; tail -40 gpurun_out/pytest_dropin.log | cut -c1-600
