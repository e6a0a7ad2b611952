# end-to-end leg: staging copies with 32-byte / 16-byte non-temporal stores and with memcpy (under gpurun)
mkdir -p gpurun_out
grep -m1 -o "avx2\|avx512f" /proc/cpuinfo | sort -u | tr '\n' ' '; echo
for nt in 1 16 1 16 0; do
  ZFB_COPY_NT=$nt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/e2e_nt$nt.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('nt=$nt e2e', round(e['value'],2), 'ms', round(e['ms_per_step'],1), 'c', round(e['ms_compress'],1), 'd', round(e['ms_decompress'],1), 'pinned', round(e['pinned']['value'],2))
"
done
