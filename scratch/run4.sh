timeout 120 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -2 > gpurun_out/s4_tests.log; cat gpurun_out/s4_tests.log
grep -q passed gpurun_out/s4_tests.log || exit 1
grep -q failed gpurun_out/s4_tests.log && exit 1
timeout 120 python scratch/small_layers.py 2>&1 | tail -12 > gpurun_out/small_layers_s4.txt
timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s4.json 2> gpurun_out/s4.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])
P
