timeout 200 python bench.py --steps 20 --warmup 3 --cpu-seconds 1 --no-extras > gpurun_out/s16_k20.json 2> gpurun_out/s16_k20.err
timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s16_k200.json 2> gpurun_out/s16_k200.err
python - <<'P'
import json
for i in ('k20','k200'):
    d=json.load(open('gpurun_out/s16_%s.json'%i)); print(i, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('clocks'))
P
