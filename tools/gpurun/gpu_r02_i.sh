# round 2: boundary kernel v2 -- suite, smoke launch list (ncu), C1-size driver run, then N-GPU scaling bench when N>1
nvidia-smi -L | head -2
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02_smoke_launches.csv') if l.startswith('"'))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
t=collections.defaultdict(float); n=collections.Counter()
for r in rows[1:]:
    t[r[ik][:70]]+=float(r[iv].replace(",","")); n[r[ik][:70]]+=1
for k,v in sorted(t.items(),key=lambda x:-x[1])[:8]: print("%10.1f us %4d  %s"%(v/1e3,n[k],k))
PY
mkdir -p /tmp/c1 && dualsphysics_multilayer_b200/bin/mkcase dambreak 170000 /tmp/c1 Case > /dev/null && dualsphysics_multilayer_b200/bin/dsph_b200 /tmp/c1/Case /tmp/c1/out -sv:none -nsteps:300 | tail -14
