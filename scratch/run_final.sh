# final measurements of the round on one GPU: tests, config B with extras, configs A / C / E, the C++ host
python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
for c in A C E; do
  timeout 300 python bench.py --config $c --steps 50 --warmup 5 --cpu-seconds 4 > gpurun_out/r2g_cfg$c.json 2> gpurun_out/r2g_cfg$c.err
done
timeout 300 python scripts/host_bench.py > gpurun_out/r2g_host.json 2> gpurun_out/r2g_host.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err
python - <<'P'
import json
for f in ('r2g_bench','r2g_cfgA','r2g_cfgC','r2g_cfgE','r2g_ref'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d.get('parity'), d.get('autotune_calls'), (d.get('gpu_eager_baseline') or {}).get('tf32_off',{}).get('ms_per_step'), (d.get('gpu_eager_baseline') or {}).get('tf32_on',{}).get('ms_per_step'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
P
tail -1 gpurun_out/r2g_host.json | cut -c1-300
