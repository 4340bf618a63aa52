# quick GPU check used while tuning the force kernels: variant agreement + timings at DB(20 M) and NN(5 M)
nvidia-smi -L
(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or nn_ring") > gpurun_out/pytest_variants.log 2>&1; tail -5 gpurun_out/pytest_variants.log
(time timeout 400 python tools/forcebench.py 20000000 8 ${VARIANTS:-6,7,8,6,7,8}) > gpurun_out/forcebench.log 2>&1; tail -8 gpurun_out/forcebench.log
if [ -n "$NN" ]; then
  (timeout 400 python tools/forcebench.py 5000000 20 1,8 nn2layer) > gpurun_out/forcebench_nn.log 2>&1
  (timeout 400 python tools/forcebench.py 5000000 20 1,8 nn2layer velgrad=2) >> gpurun_out/forcebench_nn.log 2>&1
  (timeout 400 python tools/forcebench.py 5000000 20 1,8 nn2layer velgrad=2 tvisco=3) >> gpurun_out/forcebench_nn.log 2>&1
  grep "variant" gpurun_out/forcebench_nn.log
fi
