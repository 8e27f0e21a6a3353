cp esmstereo_b200/plans/b200.txt gpurun_out/b200_old.txt
ESM_PLANS=0 ESM_AUTOTUNE=1 timeout 900 python scripts/tune_plans.py gpurun_out/b200_new.txt > gpurun_out/tune.log 2>&1; tail -3 gpurun_out/tune.log
test -s gpurun_out/b200_new.txt || exit 1
cp gpurun_out/b200_new.txt esmstereo_b200/plans/b200.txt
timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > gpurun_out/s5.json 2> gpurun_out/s5.err
python - <<'P'
import json
for f in ('gpurun_out/s5.json',):
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d.get('autotune_calls'))
P
