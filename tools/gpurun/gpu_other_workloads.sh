# the other single-GPU configs of BASELINE.json: periodic box PB(10 M) / PB(40 M) and the two-layer non-Newtonian dam break NN(5 M)
for w in "periodicbox 10000000" "periodicbox 40000000" "nn2layer 5000000"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --np $2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-reference-gpu > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  cut -c1-230 gpurun_out/bench_$1_$2.json; tail -2 gpurun_out/bench_$1_$2.err
done
