# end-to-end leg only, for a few copy-thread counts (under gpurun)
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled; nproc
for t in ${THREADS:-4 8 16}; do
  ZFB_COPY_THREADS=$t timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/e2e_$t.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('threads $t e2e', round(e['value'],2), 'ms', round(e['ms_per_step'],1), 'c', round(e['ms_compress'],1), 'd', round(e['ms_decompress'],1), 'pinned', round(e['pinned']['value'],2))
"
done
