timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s6_tests.log; cat gpurun_out/s6_tests.log
grep -q passed gpurun_out/s6_tests.log || exit 1
grep -q failed gpurun_out/s6_tests.log && exit 1
ESM_FUSE_ASSEMBLY=0 timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s6_nofuse.json 2> gpurun_out/s6_nofuse.err
timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > gpurun_out/s6.json 2> gpurun_out/s6.err
python - <<'P'
import json
for f in ('gpurun_out/s6_nofuse.json','gpurun_out/s6.json'):
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d.get('autotune_calls'))
P
