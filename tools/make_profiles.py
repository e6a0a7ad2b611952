#!/usr/bin/env python
"""Turn the outputs of tools/gpu_profile_final.sh (gpurun_out/) into the tracked evidence under profiles/:
  python tools/make_profiles.py r02
-> profiles/<tag>_ncu_full_nyx2048slab.csv, <tag>_traffic_nyx2048slab.json, <tag>_launches_bench_nyx2048slab.csv,
   <tag>_hot_lines_nyx2048slab.txt"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
rep = os.path.join(G, "prof_final_slab.ncu-rep")
full = os.path.join(P, "%s_ncu_full_nyx2048slab.csv" % tag)
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "--csv", full], capture_output=True)
shutil.copy(os.path.join(G, "launches_bench_final.csv"), os.path.join(P, "%s_launches_bench_nyx2048slab.csv" % tag))
d = {}
for r in csv.DictReader(open(full)):
    k = "encode" if "enc3" in r["kernel"] else "decode+metrics"
    d.setdefault(k, {"kernel": r["kernel"].split("(")[0].replace("void ", "")})[r["metric"]] = r["value"]
def unit_bytes(v):  # ncu prints dram bytes in G (base 10) in this summary
    return float(v) * 1e9
out = {"workload": "nyx2048-slab (256x2048x2048 f32, 8 bpv)",
       "command": "ncu --set full --clock-control none --import-source on --kernel-name regex:\"(enc3_f32_warp|dec3_f32_ws)\" -s 6 -c 2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu  (tools/gpu_profile_final.sh)",
       "algorithmic_bytes": {"encode": 5368709120, "decode+metrics": 9663676416},
       "note": "algorithmic bytes: field 4 GiB (N*s) + stream 1 GiB (8 bpv) for encode; stream + original re-read + decoded write for decode+metrics (SURVEY 8d)",
       "kernels": {}}
for k, m in d.items():
    rd, wr = unit_bytes(m["dram__bytes_read.sum"]), unit_bytes(m["dram__bytes_write.sum"])
    out["kernels"][k] = {"kernel": m["kernel"], "dram_bytes_read": rd, "dram_bytes_write": wr,
                         "duration_ms_under_ncu": float(m["gpu__time_duration.sum"]),
                         "dram_throughput_pct_of_peak": float(m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]),
                         "registers": int(m["launch__registers_per_thread"]),
                         "ipc_per_sm": float(m["sm__inst_executed.avg.per_cycle_elapsed"]),
                         "alu_pipe_pct": float(m["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]),
                         "fma_pipe_pct": float(m["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]),
                         "warp_instructions": float(m["smsp__inst_executed.sum"]),
                         "active_threads_per_instruction": float(m["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                         "traffic": rd + wr}
json.dump(out, open(os.path.join(P, "%s_traffic_nyx2048slab.json" % tag), "w"), indent=1)
hot = os.path.join(P, "%s_hot_lines_nyx2048slab.txt" % tag)
with open(hot, "w") as f:
    f.write("# Per-source-line warp instruction counts of the two float kernels at the benchmark size (256x2048x2048 f32, 8 bpv)\n"
            "# from the ncu --set full --import-source on capture behind %s_ncu_full_nyx2048slab.csv (tools/gpu_profile_final.sh), joined with the\n"
            "# cubin's line table (tools/ncu_lines.py).  columns: share of the kernel's warp instructions, count, average active threads per\n"
            "# instruction, stall samples, file:line, source\n\n" % tag)
    for pat in ("enc3_f32_warp", "dec3_f32_ws"):
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, pat, "50"], capture_output=True, text=True).stdout + "\n")
print(json.dumps(out["kernels"], indent=1))
