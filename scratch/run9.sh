timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_kernels python scripts/prof_conv.py > gpurun_out/prof_kernels.log 2>&1
for c in A C E; do
  timeout 300 python bench.py --config $c --steps 50 --warmup 5 --cpu-seconds 4 > gpurun_out/r2f_cfg$c.json 2> gpurun_out/r2f_cfg$c.err
done
timeout 300 python scripts/host_bench.py > gpurun_out/r2f_host.json 2> gpurun_out/r2f_host.err
python - <<'P'
import json
for f in ('r2f_bench','r2f_cfgA','r2f_cfgC','r2f_cfgE'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d.get('parity'), d.get('autotune_calls'))
    except Exception as e: print(f,'ERR',e)
P
tail -1 gpurun_out/r2f_host.json | cut -c1-200
