timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pf.py -x -q 2>&1 | tail -2 > gpurun_out/s4_tests.log; cat gpurun_out/s4_tests.log
grep -q passed gpurun_out/s4_tests.log || exit 1
grep -q failed gpurun_out/s4_tests.log && exit 1
timeout 120 python scratch/small_layers.py 2>&1 | tail -12 > gpurun_out/small_layers_s4.txt
ESM_TC_WIMG=0 timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s4_img0.json 2> gpurun_out/s4_img0.err
timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > gpurun_out/s4.json 2> gpurun_out/s4.err
python - <<'P'
import json
for f in ('gpurun_out/s4_img0.json','gpurun_out/s4.json'):
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'))
P
