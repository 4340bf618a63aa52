# round 2: whole GPU suite, smoke, default bench (all legs)
nvidia-smi -L
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
(time timeout 1500 python bench.py > gpurun_out/bench_default.json) 2> gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
