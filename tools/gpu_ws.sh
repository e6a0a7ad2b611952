# Warp-specialised decode kernel: parity subset + A/B against the tile kernel (ZFB_DEC_WS=0)
mkdir -p gpurun_out
(timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_ws.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ws.log); tail -5 gpurun_out/pytest_ws.log
for ws in 1 0; do
for w in ${WORKLOADS:-nyx2048-slab nyx512}; do
ZFB_DEC_WS=$ws timeout -k 10 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ws=$ws',d['config']['workload'],'value',round(d['value'],1),'enc_ms',round(d['enc_ms'],3),'dec_ms',round(d['dec_ms'],3),'frac dec',round(d['roofline']['frac'],3),'rt',round(d['roofline']['round_trip']['frac'],3))
    else: print(l.strip()[:300])
"
done
done
