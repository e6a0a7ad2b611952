# HACC 1D: bench line (no ncu), then one ncu --set full capture of the two line kernels with source info
mkdir -p gpurun_out
timeout 200 python bench.py --workload hacc1d --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('hacc1d value',round(d['value'],1),'enc_ms',round(d['enc_ms'],3),'dec_ms',round(d['dec_ms'],3),'rt',round(d['roofline']['round_trip']['frac'],3))
    else: print(l.strip()[:200])
"
[ -n "$NO_NCU" ] && exit 0
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"enc1_f32|dec1_f32" -s 4 -c 2 -o gpurun_out/prof_hacc1d_r2 -f \
  python bench.py --workload hacc1d --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_hacc1d_r2.log 2>&1
tail -1 gpurun_out/ncu_hacc1d_r2.log
