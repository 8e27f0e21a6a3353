timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pf.py tests/test_fullsize_golden.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s11_tests.log; cat gpurun_out/s11_tests.log
grep -q passed gpurun_out/s11_tests.log || exit 1
grep -q failed gpurun_out/s11_tests.log && exit 1
ESM_TC_TAILHELP=0 timeout 200 python scratch/small_layers.py 2>&1 | head -5 > gpurun_out/sl_h0.txt
timeout 200 python scratch/small_layers.py 2>&1 | head -5 > gpurun_out/sl_h1.txt
ESM_TC_TAILHELP=0 timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s11_h0.json 2> gpurun_out/s11_h0.err
timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s11_h1.json 2> gpurun_out/s11_h1.err
python - <<'P'
import json
for f in ('gpurun_out/s11_h0.json','gpurun_out/s11_h1.json'):
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
P
