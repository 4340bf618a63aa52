# round 2: all tests that need 2 GPUs + the N-GPU bench of this box (gpurun --gpus N)
N=${N:-2}
nvidia-smi -L | wc -l
(time timeout 1500 python -m pytest tests -m gpu -x -q -k "two_slabs or slab_run or nn3 or bound") > gpurun_out/pytest_mgpu.log 2>&1; tail -5 gpurun_out/pytest_mgpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err; cut -c1-260 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
