# round 2, first GPU job: parity suite, v9 vs v10 at 20 M, full default bench (new measurement contract)
nvidia-smi -L
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
(time timeout 400 python tools/forcebench.py 20000000 8 9,10,9,10) > gpurun_out/forcebench.log 2>&1; tail -6 gpurun_out/forcebench.log
(time timeout 1500 python bench.py > gpurun_out/bench_default.json) 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-600; tail -4 gpurun_out/bench_default.err
