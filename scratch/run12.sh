timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s12_tests.log; cat gpurun_out/s12_tests.log
grep -q passed gpurun_out/s12_tests.log || exit 1
grep -q failed gpurun_out/s12_tests.log && exit 1
timeout 200 python bench.py --steps 100 --warmup 5 --cpu-seconds 1 > gpurun_out/s12.json 2> gpurun_out/s12.err
python - <<'P'
import json
for f in ('gpurun_out/s12.json',):
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d.get('autotune_calls'))
P
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
