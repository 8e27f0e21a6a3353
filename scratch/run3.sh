timeout 120 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pf.py -x -q 2>&1 | tail -3 > gpurun_out/s3_tests3.log; cat gpurun_out/s3_tests3.log
grep -q passed gpurun_out/s3_tests3.log || exit 1
grep -q failed gpurun_out/s3_tests3.log && exit 1
ESM_TC_KHK=1 timeout 120 python scratch/small_layers.py 2>&1 | tail -12 > gpurun_out/small_layers_khk1.txt
ESM_TC_PROFILE=1 timeout 120 python scratch/small_layers.py prof 2>&1 | grep -v "warp  0" | head -80 > gpurun_out/small_layers_prof3.txt
ESM_TC_KHK=1 timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > gpurun_out/s3_khk1.json 2> gpurun_out/s3_khk1.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s3_khk1.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'))
    except Exception as e: print(f, 'ERR', e)
P
