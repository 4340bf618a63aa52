# ncu evidence for profiles/: launch list of a short bench.py run + one --set full capture of the production fluid force kernel
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-reference-gpu"
timeout 600 $B > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench20m.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ker_forces_fluid_v5 -s 3 -c 1 -f -o gpurun_out/forces_v9_20m $B > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
