# round 2: N-GPU weak-scaling bench with the native transport (gpurun --gpus N)
N=${N:-8}
nvidia-smi -L | wc -l
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
