# round 2, final single-GPU evidence: the default bench.py run (as the driver runs it), the launch list of a short run, smoke()
nvidia-smi -L | head -1
(time timeout 1500 python bench.py > gpurun_out/bench20m_default.json) 2> gpurun_out/bench20m_default.err; cut -c1-300 gpurun_out/bench20m_default.json; tail -4 gpurun_out/bench20m_default.err
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-reference-gpu --no-ncu --developed-step 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench20m.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
