cp esmstereo_b200/plans/b200.txt /tmp/b200_old.txt
ESM_PLANS=0 ESM_AUTOTUNE=1 timeout 900 python scripts/tune_plans.py gpurun_out/b200_new.txt > gpurun_out/tune.log 2>&1; tail -1 gpurun_out/tune.log
test -s gpurun_out/b200_new.txt || exit 1
python - <<'P'
old=[l for l in open('/tmp/b200_old.txt') if not l.startswith('#') and ':' in l]
newl=open('gpurun_out/b200_new.txt').read().splitlines(keepends=True)
hdr=[l for l in newl if l.startswith('#')]
new=[l for l in newl if not l.startswith('#') and ':' in l]
key=lambda l:l.split(':')[0].strip()
nk={key(l) for l in new}
extra=[l for l in old if key(l) not in nk]
od={key(l):l for l in old}
print(len(old),len(new),len(extra),'engine changed for',sum(1 for l in new if key(l) in od and l.split(':')[1].split()[0]!=od[key(l)].split(':')[1].split()[0]))
open('esmstereo_b200/plans/b200.txt','w').write(''.join(hdr+new+extra))
open('gpurun_out/b200_merged.txt','w').write(''.join(hdr+new+extra))
P
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/s22_tests.log; cat gpurun_out/s22_tests.log
for i in 1 2; do timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s22_$i.json 2> gpurun_out/s22_$i.err; done
cp /tmp/b200_old.txt esmstereo_b200/plans/b200.txt
timeout 200 python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/s22_oldplans.json 2> gpurun_out/s22_oldplans.err
python - <<'P'
import json
for k in ('1','2','oldplans'):
    d=json.load(open('gpurun_out/s22_%s.json'%k)); print(k, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('autotune_calls'))
P
