# round 2: slab tests (peer-memory transport) + N-GPU bench with phase profile
N=${N:-2}
nvidia-smi -L | wc -l
[ "$N" = 2 ] && (time timeout 1500 python -m pytest tests -m gpu -q -k "slab") > gpurun_out/pytest_slab.log 2>&1; tail -15 gpurun_out/pytest_slab.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err; cut -c1-260 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n$N.json')); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['timers_ms_per_step']); print(d.get('mgpu_phases_ms_per_step'))"
