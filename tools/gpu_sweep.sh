# rate sweeps for the tables in DESIGN.md (run on the GPU box; writes gpurun_out/sweep.jsonl)
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for spec in "nyx512 4" "nyx512 8" "nyx512 16" "nyx2048-slab 4" "nyx2048-slab 16" "hacc1d 8" "hacc1d 16" "vpic1024-f64 1" "vpic1024-f64 4" "vpic1024-f64 8" "vpic1024-f64 16" "vpic1024-f64 32"; do
  set -- $spec
  timeout -k 10 300 python bench.py --workload $1 --rate $2 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep '^{' >> gpurun_out/sweep.jsonl
done
# coder best / worst cases, HACC velocities, unaligned (wra=0) 1D slots
for spec in "nyx512 8 noise 1" "nyx512 8 zero 1" "hacc1d 8 velocity 1" "hacc1d 8 uniform 0" "hacc1d 16 uniform 0" "hacc1d 8 noise 1" "vpic1024-f64 8 noise 1" "vpic1024-f64 8 zero 1"; do
  set -- $spec
  timeout -k 10 300 python bench.py --workload $1 --rate $2 --field $3 --wra $4 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep '^{' >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    d=json.loads(l); c=d["config"]
    print(c["workload"], c.get("field"), "wra", c.get("wra"), "rate", c["rate_bpv"], "maxbits", c["maxbits"], "GB/s", round(d["value"],1), "enc_ms", round(d["enc_ms"],3), "dec_ms", round(d["dec_ms"],3), "rt_frac", round(d["roofline"]["round_trip"]["frac"],3), "psnr", round(d["metrics_check"]["psnr"],2))
PY
