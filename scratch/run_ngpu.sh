# bench on N GPUs of one box: bash scratch/run_ngpu.sh N   (under gpurun --gpus N)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/r2g_n$N.json 2> gpurun_out/r2g_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2g_n$N.json').read().strip().splitlines()[-1]); print($N, d['value'], d['ms_per_step'], d.get('rank_spread'), d['e2e']['value'])
P
