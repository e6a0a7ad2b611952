mkdir -p gpurun_out
cmd="python bench.py --workload nyx512 --steps 1 --warmup 3 --no-e2e --no-cpu --hist"
timeout 300 $cmd > gpurun_out/plain_hist.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"dec3_f32_tile" -s 3 -c 1 -o gpurun_out/prof_hist -f $cmd > gpurun_out/ncu_hist.log 2>&1
tail -2 gpurun_out/ncu_hist.log
