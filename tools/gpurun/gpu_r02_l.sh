# round 2: mDBC against the reference dumps + driver run of an mDBC case against the reference CPU solver
nvidia-smi -L | head -1
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "mdbc") > gpurun_out/pytest_mdbc.log 2>&1; tail -14 gpurun_out/pytest_mdbc.log | cut -c1-700
