"""PAT compatibility of the CSV the plugins produce: Foresight's Python workflow reads CBench's
metrics_.csv (CBench/main.cpp:125-129 header, :259-260,:340 row prefix and metric columns, :414-432 throughput / ratio
and file write) in Analysis/pat/nyx/cinema.py:17-54 and Analysis/pat/hacc/cinema.py:47-80.

  * GPU test: cbench_mini (the main.cpp scalar loop around the plugin classes) produces the rows for two compressors x
    the six NYX scalars of the toy file; PAT's parsing -- restated below line by line -- must accept them.
  * CPU test: the committed sample of exactly that output (tests/golden/metrics_sample.csv, made by
    tools/make_metrics_sample.py on a B200) goes through the REAL pat.nyx.cinema where /root/reference exists
    (matplotlib is stubbed: the image is not needed to parse a CSV) and through the restatement; both agree."""
import csv
import io
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(os.path.dirname(__file__), "golden")
H5 = os.path.join(GOLD, "z255_32.h5")
SAMPLE = os.path.join(GOLD, "metrics_sample.csv")
PAT = "/root/reference/Analysis"
NYX_FIELDS = ["baryon_density", "dark_matter_density", "temperature", "velocity_x", "velocity_y", "velocity_z"]
COMPRESSORS = [("zfp_gpu", "8", "zfp_gpu_8_"), ("zfp", "abs=0.01", "zfp_abs_.01_")]
HEADER = ["Compressor_field__params", "name", "absolute_error", "relative_error", "mse", "psnr", "minmax",
          "Compression Throughput(MB/s)", "DeCompression Throughput(MB/s)", "Compression Ratio"]


def build_metrics_csv(pkg, workdir):
    """metrics_.csv as main.cpp accumulates it: one header, then one row per (compressor, scalar)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cbench_plugin import build_mini

    exe = build_mini()
    raw = open(H5, "rb").read()
    text = None
    for cname, arg, prefix in COMPRESSORS:
        for field in NYX_FIELDS:
            d = pkg.h5_dataset(H5, "/native_fields/" + field)
            path = os.path.join(workdir, field + ".f32")
            open(path, "wb").write(raw[d["offset"]:d["offset"] + d["bytes"]])
            r = subprocess.run([exe, path, "32", "32", "32", arg, "--field", field, "--name", cname, "--prefix", prefix],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
            lines = r.stdout.splitlines()
            if text is None:
                text = lines[0] + "\n"
            assert lines[0] + "\n" == text.splitlines()[0] + "\n"
            text += lines[1] + "\n"
    return text


def pat_nyx_prepare(csv_text):
    """Analysis/pat/nyx/cinema.py:27-52 (prepare_cinema), on text instead of files."""
    rows = []
    reader = csv.reader(io.StringIO(csv_text))
    row = next(reader)
    row += ["FILE_SimStats_Pk", "FILE_lya_all_axes_x_Pk", "FILE_lya_all_axes_y_Pk", "FILE_lya_all_axes_z_Pk"]
    rows.append(row)
    values = ["sim_stats_rhob.png", "sim_stats_rhodm.png", "sim_stats_temp.png", "sim_stats_velmag.png",
              "sim_stats_velmag.png", "sim_stats_vz.png"]
    count = 0
    for row in reader:
        row += [values[count], "lya_all_axes_x.png", "lya_all_axes_y.png", "lya_all_axes_z.png"]
        rows.append(row)
        count += 1
        if count == 6:
            count = 0
    return rows


def check_rows(rows):
    assert [c.strip() for c in rows[0][:10]] == HEADER
    assert len(rows) == 1 + len(COMPRESSORS) * len(NYX_FIELDS)
    k = 1
    for cname, arg, prefix in COMPRESSORS:
        for j, field in enumerate(NYX_FIELDS):
            row = rows[k]
            assert len(row) == len(rows[0]), row          # a stray comma in a name or parameter would shift the columns
            params = arg.replace("=", "") if "=" in arg else "bits" + arg   # getParamsInfo(): key + value
            assert row[0] == "%s_%s__%s" % (cname, field, params)
            assert row[1][1:] == prefix                    # pat/hacc/cinema.py:58: prefix = row[1][1:]
            vals = [float(t) for t in row[2:10]]           # every metric / throughput / ratio column is a number
            assert all(np.isfinite(v) for v in vals)
            assert vals[7] > 1.0                           # compression ratio
            # the per-scalar image column cycles with the six NYX scalars (cinema.py:40-52)
            assert row[10] == ["sim_stats_rhob.png", "sim_stats_rhodm.png", "sim_stats_temp.png", "sim_stats_velmag.png",
                               "sim_stats_velmag.png", "sim_stats_vz.png"][j]
            k += 1


@pytest.mark.gpu
def test_metrics_csv_is_what_pat_parses(pkg, tmp_path):
    text = build_metrics_csv(pkg, str(tmp_path))
    rows = pat_nyx_prepare(text)
    check_rows(rows)
    if os.path.exists(SAMPLE):   # same shape as the committed sample (the numbers in the throughput columns differ)
        old = list(csv.reader(open(SAMPLE)))
        new = list(csv.reader(io.StringIO(text)))
        assert [r[:7] for r in old] == [r[:7] for r in new]


def test_committed_sample_through_pat(tmp_path):
    if not os.path.exists(SAMPLE):
        pytest.skip("tests/golden/metrics_sample.csv has not been generated yet (tools/make_metrics_sample.py, needs a GPU)")
    text = open(SAMPLE).read()
    rows = pat_nyx_prepare(text)
    check_rows(rows)
    if not os.path.isdir(PAT):
        return
    # the real thing: pat.nyx.cinema.nyx_cinema.prepare_cinema on files laid out as the workflow does
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.colors", "matplotlib.cm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.path.insert(0, PAT)
    # pat/nyx/cinema.py runs its whole workflow at import time (argparse + create_plots at module level): only the class
    # definitions above that tail are executed here, from the file where it lies
    src = open(os.path.join(PAT, "pat", "nyx", "cinema.py")).read()
    src = src[:src.index("# Parse Input")]
    pc = types.ModuleType("pat_nyx_cinema_classes")
    try:
        exec(compile(src, os.path.join(PAT, "pat", "nyx", "cinema.py"), "exec"), pc.__dict__)
    except Exception as e:  # some other dependency of the plotting helpers is missing here
        pytest.skip("pat.nyx.cinema not importable: %r" % (e,))
    home = str(tmp_path) + "/"
    os.makedirs(home + "wf/cbench")
    os.makedirs(home + "wf/cinema")
    open(home + "wf/cbench/metrics_.csv", "w").write(text)
    import json
    cfg = {"project-home": home, "wflow-path": "wf", "input": {}, "cbench": {"output": {"metrics-file": "metrics_"}}}
    jf = home + "wf.json"
    json.dump(cfg, open(jf, "w"))
    wf = pc.nyx_cinema(jf)
    wf.prepare_cinema()
    got = list(csv.reader(open(home + "wf/cinema/data.csv")))
    assert got == rows
