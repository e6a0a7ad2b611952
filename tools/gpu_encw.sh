# warp-autonomous encode kernel: parity subset + A/B against the tile kernel (ZFB_ENC_WARP=0)
mkdir -p gpurun_out
(timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/pytest_encw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_encw.log); tail -3 gpurun_out/pytest_encw.log
for ew in 1 0 1 0; do
ZFB_ENC_WARP=$ew timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('enc_warp=$ew','value',round(d['value'],1),'enc_ms',round(d['enc_ms'],3),'dec_ms',round(d['dec_ms'],3),'enc frac',round(d['roofline']['encode']['frac'],3),'rt',round(d['roofline']['round_trip']['frac'],3))
    else: print(l.strip()[:300])
"
done
