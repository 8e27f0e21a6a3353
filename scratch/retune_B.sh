# re-time the engines for the KITTI-shaped forward only and A/B the result against the pinned plans (diagnostics)
ESM_PLANS=0 ESM_AUTOTUNE=1 python scripts/tune_plans.py gpurun_out/b200_B.txt --only 0 2>&1 | tail -2
ESM_PLANS=0 ESM_AUTOTUNE=1 python scripts/tune_plans.py gpurun_out/b200_B2.txt --only 0 2>&1 | tail -1
python bench.py --steps 200 --warmup 5 --cpu-seconds 1 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pinned', d['ms_per_step'])"
cp gpurun_out/b200_B.txt esmstereo_b200/plans/b200_zB.txt
python bench.py --steps 200 --warmup 5 --cpu-seconds 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('retuned', d['ms_per_step'], d['parity'])"
rm esmstereo_b200/plans/b200_zB.txt
