# round 2: native NCCL transport (dsph_mgpu_step) at 2 GPUs: slab parity tests, then bench.py N=2 native vs python exchange
N=${N:-2}
nvidia-smi -L
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_slabs") > gpurun_out/pytest_mgpu.log 2>&1; tail -6 gpurun_out/pytest_mgpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 10 --warmup 3 --developed-step 0 --no-e2e --profile-mgpu > gpurun_out/bench_n${N}_native.json) 2> gpurun_out/bench_n${N}_native.err; cut -c1-400 gpurun_out/bench_n${N}_native.json; tail -5 gpurun_out/bench_n${N}_native.err
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $N --steps 10 --warmup 3 --developed-step 0 --no-e2e --profile-mgpu --python-mgpu > gpurun_out/bench_n${N}_python.json) 2> gpurun_out/bench_n${N}_python.err; cut -c1-400 gpurun_out/bench_n${N}_python.json; tail -5 gpurun_out/bench_n${N}_python.err
