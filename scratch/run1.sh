python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; tail -3 gpurun_out/s3_tests.log
for pdl in 0 62 63; do
  ESM_PDL=$pdl python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s3_pdl$pdl.json 2> gpurun_out/s3_pdl$pdl.err
done
ESM_PDL=0 ESM_SMLAYER=0 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s3_pdl0_sm0.json 2> gpurun_out/s3_pdl0_sm0.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s3_pdl*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
P
