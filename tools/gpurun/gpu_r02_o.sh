# round 2 evidence for profiles/: launch list + full ncu capture of the production fluid kernel at DB(20 M), the other single-GPU
# configurations of BASELINE.json (PB 10 M / 40 M, NN 5 M), a compute-sanitizer pass of small parity cases
nvidia-smi -L | head -1
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-reference-gpu --no-ncu --developed-step 0"
timeout 600 $B > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench20m.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ker_forces_fluid_v10 -s 3 -c 1 -f -o gpurun_out/forces_v10_20m $B > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/forces_v10_20m.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_forces_v10_20m.csv 2>/dev/null
for w in "periodicbox 10000000" "periodicbox 40000000" "nn2layer 5000000"; do
  set -- $w
  timeout 900 python bench.py --workload $1 --np $2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-reference-gpu > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  cut -c1-230 gpurun_out/bench_$1_$2.json; tail -2 gpurun_out/bench_$1_$2.err
done
(time timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vs_oracle_steps and (dambreak-300 or dambreak-2500 or periodicbox-343 or nn2layer-1500)") > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -8 gpurun_out/r02_sanitizer_racecheck.log
(time timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gauges.py -m gpu -x -q -k "(vs_oracle_steps and (dambreak-2500 or periodicbox-343 or nn2layer-1500)) or gauges_match or (stages_vs_reference and mdbc)") > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -8 gpurun_out/r02_sanitizer_memcheck.log
ls -la gpurun_out | head -5
