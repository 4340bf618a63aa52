nvidia-smi -L | head -1
timeout 600 python tools/mdbc_variants.py 0 1 > gpurun_out/mdbc_variants.log 2>&1; cat gpurun_out/mdbc_variants.log | cut -c1-250
