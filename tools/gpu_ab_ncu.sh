# A/B instruction counts: ncu (few metrics) on the nyx512 workload for alternative builds libzfb_<tag>.so
mkdir -p gpurun_out
cd vizaly-foresight_b200; cp libzfb.so libzfb_orig.so; cd ..
for tag in "$@"; do
  cp vizaly-foresight_b200/libzfb_$tag.so vizaly-foresight_b200/libzfb.so
  timeout 120 python bench.py --workload nyx512 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1 || echo "plain run failed for $tag"
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,launch__registers_per_thread,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,sass__inst_executed_register_spilling \
    --clock-control none --kernel-name regex:"(enc3|dec3|enc1|dec1)_" -c 2 --csv --log-file gpurun_out/ab_$tag.csv python bench.py --workload ${WL:-nyx512} --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ab_$tag.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print('$tag', d['Kernel Name'][:28], d['Metric Name'], d['Metric Value'])
PY
done
cp vizaly-foresight_b200/libzfb_orig.so vizaly-foresight_b200/libzfb.so
