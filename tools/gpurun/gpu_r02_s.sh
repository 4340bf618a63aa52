# round 2: periodic box over a ring of x-slabs: 2-GPU tests + PB(2 x 10 M) bench
N=${N:-2}
(time timeout 900 python -m pytest tests -m gpu -q -k "two_slabs and periodicbox") > gpurun_out/pytest_pslab.log 2>&1; tail -5 gpurun_out/pytest_pslab.log; grep "MGPU_CHECK\|Error\|error" gpurun_out/pytest_pslab.log | cut -c1-300 | head -20
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --workload periodicbox --particles 10000000 --gpus $N --steps 10 --warmup 3 --profile-mgpu > gpurun_out/bench_pb_n$N.json) 2> gpurun_out/bench_pb_n$N.err; cut -c1-300 gpurun_out/bench_pb_n$N.json; tail -5 gpurun_out/bench_pb_n$N.err
