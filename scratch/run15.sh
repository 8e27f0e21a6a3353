timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s15_w5.json 2> gpurun_out/s15_w5.err
timeout 200 python bench.py --steps 200 --warmup 3000 --cpu-seconds 1 --no-extras > gpurun_out/s15_w3000.json 2> gpurun_out/s15_w3000.err
timeout 200 python bench.py --steps 2000 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s15_k2000.json 2> gpurun_out/s15_k2000.err
python - <<'P'
import json
for i in ('w5','w3000','k2000'):
    d=json.load(open('gpurun_out/s15_%s.json'%i)); print(i, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('clocks'))
P
