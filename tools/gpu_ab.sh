# A/B harness: run the bench with alternative builds of libzfb (libzfb_<tag>.so) swapped in
mkdir -p gpurun_out
cd vizaly-foresight_b200; cp libzfb.so libzfb_orig.so; cd ..
for tag in "$@"; do
  cp vizaly-foresight_b200/libzfb_$tag.so vizaly-foresight_b200/libzfb.so
  timeout -k 10 300 python bench.py ${BENCH_ARGS:-} --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$tag','value',round(d['value'],1),'enc_ms',round(d['enc_ms'],3),'dec_ms',round(d['dec_ms'],3),'rt',round(d['roofline']['round_trip']['frac'],3),'psnr',d['metrics_check']['psnr'],'abs',d['metrics_check']['absolute_error'],'mse',d['metrics_check']['mse'])
    else: print(l.strip()[:200])
"
done
cp vizaly-foresight_b200/libzfb_orig.so vizaly-foresight_b200/libzfb.so
