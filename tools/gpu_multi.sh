# multi-GPU check (under gpurun --gpus N): bench at N ranks, weak and strong, with and without the fused histogram
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('$1', d['config']['workload'], d['scaling'], 'n_gpus', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'enc', round(d['enc_ms'],3), 'dec', round(d['dec_ms'],3), 'e2e', e.get('value'), 'hist', d['config'].get('histogram'), d['metrics_check'])
"; }
run --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/multi_weak.err | tee gpurun_out/multi_weak_$N.json | show weak
run --steps 10 --warmup 3 --no-cpu --no-e2e --hist 2>gpurun_out/multi_hist.err | tee gpurun_out/multi_hist_$N.json | show hist
run --steps 5 --warmup 3 --no-cpu --no-e2e --scaling strong 2>gpurun_out/multi_strong.err | tee gpurun_out/multi_strong_$N.json | show strong
tail -3 gpurun_out/multi_weak.err gpurun_out/multi_hist.err gpurun_out/multi_strong.err | cut -c1-300
