mkdir -p gpurun_out
(timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-nyx2048-slab}; do
timeout -k 10 400 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'],'value',round(d['value'],1),'enc_ms',round(d['enc_ms'],3),'dec_ms',round(d['dec_ms'],3),'frac dec',round(d['roofline']['frac'],3),'enc',round(d['roofline']['encode']['frac'],3),'rt',round(d['roofline']['round_trip']['frac'],3),d['clocks'])
    else: print(l.strip()[:300])
"
done
