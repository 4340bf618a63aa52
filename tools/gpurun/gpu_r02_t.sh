# round 2, final 8-GPU numbers: DB(8 x 20 M) and PB(8 x 10 M)
N=${N:-8}
nvidia-smi -L | wc -l
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err; cut -c1-200 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
(time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 bench.py --workload periodicbox --particles 10000000 --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_pb_n$N.json) 2> gpurun_out/bench_pb_n$N.err; cut -c1-200 gpurun_out/bench_pb_n$N.json; tail -3 gpurun_out/bench_pb_n$N.err
