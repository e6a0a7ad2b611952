# ncu --set full capture of the float/3D tile kernels (one launch each, after warm-up) -> gpurun_out/prof_<tag>.ncu-rep
# usage (under gpurun): bash tools/gpu_prof.sh <tag> [workload] [kernel regex]
tag=${1:-dev}; wl=${2:-nyx512}; pat=${3:-"(enc3_f32_warp|dec3_f32_ws)"}
mkdir -p gpurun_out
cmd="python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu"
timeout 300 $cmd > gpurun_out/plain_$tag.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"$pat" -s ${SKIP:-6} -c ${COUNT:-2} -o gpurun_out/prof_$tag -f $cmd > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
grep -h '^{' gpurun_out/plain_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'],'value',round(d['value'],1),'enc_ms',round(d['enc_ms'],3),'dec_ms',round(d['dec_ms'],3))"
