nvidia-smi -L
(time timeout 600 python tools/parity_floors.py) > gpurun_out/parity_floors.log 2>&1; cat gpurun_out/parity_floors.log | cut -c1-260
