cd vizaly-foresight_b200; cp libzfb.so libzfb_orig.so; cd ..
for tag in base new base new; do
  cp vizaly-foresight_b200/libzfb_$tag.so vizaly-foresight_b200/libzfb.so
  timeout 200 python bench.py --workload vpic1024-f64 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value'],1), 'enc', round(d['enc_ms'],3), 'dec', round(d['dec_ms'],3))
"
done
cp vizaly-foresight_b200/libzfb_orig.so vizaly-foresight_b200/libzfb.so
