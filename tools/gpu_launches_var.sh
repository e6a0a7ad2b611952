# launch list (time, DRAM bytes, instructions) of one variable-rate step; WL / MODE from the environment
mkdir -p gpurun_out
WL=${WL:-nyx512}; MODE=${MODE:-abs:0.01}
timeout 120 python bench.py --workload $WL --mode $MODE --steps 2 --warmup 3 > gpurun_out/plain_var.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_var.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,launch__registers_per_thread \
  --clock-control none --kernel-name regex:"var|scan_|zero_stream|finalize" -c 16 --csv --log-file gpurun_out/launches_var.csv python bench.py --workload $WL --mode $MODE --steps 2 --warmup 3 > gpurun_out/ncu_var.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_var.csv")) if len(r)>10]
h=rows[0]; agg=collections.OrderedDict()
for r in rows[1:]:
    d=dict(zip(h,r)); k=d["Kernel Name"][:34]; agg.setdefault(k,{}).setdefault(d["Metric Name"].split("__")[1][:22],[]).append(float(d["Metric Value"].replace(",","")))
for k,v in agg.items():
    print(k, {m:round(sum(x)/len(x),1) for m,x in v.items()})
PY
