# variable-rate (fixed accuracy / precision) sweep; writes gpurun_out/var_sweep.jsonl
mkdir -p gpurun_out; : > gpurun_out/var_sweep.jsonl
for spec in ${SPECS:-"nyx512 abs:0.01" "nyx512 abs:1.0" "nyx512 rel:16" "hacc1d abs:0.003" "nyx2048-slab abs:0.1" "vpic1024-f64 abs:0.001"}; do
  set -- $spec
  timeout -k 10 300 python bench.py --workload $1 --mode $2 --steps 3 --warmup 3 2>&1 | grep '^{\|Error\|error' >> gpurun_out/var_sweep.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/var_sweep.jsonl"):
    if not l.startswith("{"): print(l.strip()[:300]); continue
    d=json.loads(l); c=d["config"]
    print(c["workload"], c["mode"], "bpv", round(c["rate_bpv"],3), "GB/s", round(d["value"],1), "enc_ms", round(d["enc_ms"],3), "dec_ms", round(d["dec_ms"],3), "rt_frac", round(d["roofline"]["round_trip"]["frac"],3), "maxabs", d["metrics_check"]["absolute_error"], "psnr", round(d["metrics_check"]["psnr"],2), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"],1) if d.get("e2e") and d["e2e"].get("value") else None)
PY
