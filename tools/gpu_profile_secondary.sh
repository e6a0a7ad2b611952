# ncu --set full captures of the secondary kernel families (1D float, 3D double, variable rate); reports in gpurun_out/
mkdir -p gpurun_out
run() {  # name, regex, bench args...
  name=$1; regex=$2; shift 2
  timeout 200 python bench.py "$@" --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1 || { echo "plain run failed: $name"; return; }
  timeout 500 ncu --set full --clock-control none --kernel-name regex:"$regex" -s 4 -c 4 -o gpurun_out/prof_sec_$name -f \
    python bench.py "$@" --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_sec_$name.log 2>&1
  tail -1 gpurun_out/ncu_sec_$name.log
}
run hacc1d "enc1_f32|dec1_f32" --workload hacc1d
run f64 "enc3_f64|dec3_f64|metrics_kernel" --workload vpic1024-f64
