# multi-GPU check (gpurun --gpus N): the 2-rank slab tests + bench.py at N ranks
N=${N:-2}
nvidia-smi -L
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_slabs") > gpurun_out/pytest_mgpu.log 2>&1; tail -4 gpurun_out/pytest_mgpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err; cut -c1-1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
