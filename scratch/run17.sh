for ms in 100 500 2000; do
ESM_CLOCK_MS=$ms timeout 200 python bench.py --steps 1000 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s17_$ms.json 2> gpurun_out/s17_$ms.err
done
python - <<'P'
import json
for i in ('100','500','2000'):
    d=json.load(open('gpurun_out/s17_%s.json'%i)); print(i, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('clocks'))
P
