mkdir -p gpurun_out
(timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
(timeout -k 10 400 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench1.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench1.log)
python - <<'PY'
import json
for l in open("gpurun_out/bench1.log"):
    if l.startswith("{"):
        d=json.loads(l); print("value",round(d["value"],1),"enc_ms",round(d["enc_ms"],3),"dec_ms",round(d["dec_ms"],3),"frac dec",round(d["roofline"]["frac"],3),"enc",round(d["roofline"]["encode"]["frac"],3),"rt",round(d["roofline"]["round_trip"]["frac"],3),d["clocks"])
    else: print(l.strip()[:300])
PY
