# under gpurun --gpus 8: PCIe / host-memory ceiling at 1, 2, 4, 8 concurrent ranks, then the bench at 8 ranks
mkdir -p gpurun_out; : > gpurun_out/pcie_probe.jsonl
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 tools/pcie_probe.py 2>/dev/null | grep '^{' | tee -a gpurun_out/pcie_probe.jsonl
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench8.err | grep '^{' | tee gpurun_out/bench8.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); e=d.get('e2e') or {}
    print('n_gpus', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', e.get('value'), 'pinned', (e.get('pinned') or {}).get('value'))
"
