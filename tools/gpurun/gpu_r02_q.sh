# round 2: suite after the store-condition fix + short bench
nvidia-smi -L | head -1
(time timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-reference-gpu --no-ncu --developed-step 0 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; cut -c1-200 gpurun_out/bench_short.json
python -c "
import json; d=json.load(open('gpurun_out/bench_short.json')); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['timers_ms_per_step'])"
