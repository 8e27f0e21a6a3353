python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/s3_tests2.log 2>&1; tail -3 gpurun_out/s3_tests2.log
ESM_TC_EPI=8 python scratch/small_layers.py > gpurun_out/small_layers_epi8.txt 2>&1
ESM_TC_EPI=16 python scratch/small_layers.py > gpurun_out/small_layers_epi16.txt 2>&1
for e in 8 16; do
  ESM_TC_EPI=$e python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s3_epi$e.json 2> gpurun_out/s3_epi$e.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s3_epi*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
P
