# round 2: the whole -m gpu suite on a box with N GPUs (the 2-GPU tests run) + the N-GPU bench + a short 1-GPU bench
N=${N:-2}
nvidia-smi -L | wc -l
(time timeout 1800 python -m pytest tests -m gpu -q) > gpurun_out/pytest_mgpu.log 2>&1; tail -5 gpurun_out/pytest_mgpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err; cut -c1-260 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-reference-gpu --no-ncu --developed-step 0 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python -c "
import json
for f in ('gpurun_out/bench_n$N.json','gpurun_out/bench_short.json'):
  d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['timers_ms_per_step'])"
