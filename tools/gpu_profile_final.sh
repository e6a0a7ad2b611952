# Final evidence for profiles/: default bench (no ncu), the reference arm, launch list of the same command, one --set full
# capture of the two tile kernels at the benchmark size.  Outputs in gpurun_out/.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || { echo bench failed; tail -5 gpurun_out/bench_default.err; exit 1; }
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench_default.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_launches.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"zfb|finalize|tile_kernel|metrics_kernel|combine" -c 400 --csv \
  --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_full.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"(enc3_f32_warp|dec3_f32_ws)" -s 6 -c 2 -o gpurun_out/prof_final_slab -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_default.json") if l.startswith("{")][-1])
e=d["e2e"]
print("value", round(d["value"],1), "enc_ms", round(d["enc_ms"],3), "dec_ms", round(d["dec_ms"],3), "frac", round(d["roofline"]["frac"],3), "rt", round(d["roofline"]["round_trip"]["frac"],3), "e2e", round(e["value"],2), "pinned", round(e["pinned"]["value"],2), "cpu", d["cpu_baseline"]["value"], d["clocks"])
r=json.loads([l for l in open("gpurun_out/bench_ref.json") if l.startswith("{")][-1])
print("reference arm", r["value"], r["unit"], r["cpu_baseline"]["cores"], "same config keys:", sorted(r["config"]) == sorted(d["config"]), [k for k in d["config"] if d["config"][k] != r["config"].get(k)])
PY
