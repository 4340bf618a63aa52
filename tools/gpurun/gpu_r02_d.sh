# round 2: driver tests (PART output set, pieces, excluded particles, -mgpu:2 slab run) at 2 GPUs + full gpu suite
nvidia-smi -L
(time timeout 900 python -m pytest tests/test_validation.py -m gpu -x -q -s -k "driver or restart") > gpurun_out/pytest_driver.log 2>&1; tail -60 gpurun_out/pytest_driver.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
