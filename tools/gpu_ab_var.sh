# A/B of alternative builds (libzfb_<tag>.so) on a variable-rate workload; WL / MODE from the environment
cd vizaly-foresight_b200; cp libzfb.so libzfb_orig.so; cd ..
for tag in "$@"; do
  cp vizaly-foresight_b200/libzfb_$tag.so vizaly-foresight_b200/libzfb.so
  timeout -k 10 300 python bench.py --workload ${WL:-nyx512} --mode ${MODE:-abs:0.01} --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$tag','value',round(d['value'],1),'enc_ms',round(d['enc_ms'],3),'dec_ms',round(d['dec_ms'],3))
"
done
cp vizaly-foresight_b200/libzfb_orig.so vizaly-foresight_b200/libzfb.so
