nvidia-smi -L
(time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree") > gpurun_out/pytest_variants.log 2>&1; tail -5 gpurun_out/pytest_variants.log
(time timeout 400 python tools/forcebench.py 20000000 8 8,9,8,9) > gpurun_out/forcebench.log 2>&1; tail -6 gpurun_out/forcebench.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
(time timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
(time timeout 500 python bench.py > gpurun_out/bench_default.json) 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-400; tail -4 gpurun_out/bench_default.err
