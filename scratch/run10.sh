ESM_PW_SPLIT=0 timeout 200 python scratch/small_layers.py 2>&1 | tail -14 > gpurun_out/sl_a.txt
ESM_TC_KHK48=1 timeout 200 python scratch/small_layers.py 2>&1 | tail -14 > gpurun_out/sl_b.txt
