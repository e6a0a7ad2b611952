#!/usr/bin/env python
"""Regenerates tests/golden/metrics_sample.csv: the metrics CSV of two compressors x the six NYX toy scalars through
cbench_mini (needs a GPU).  Under gpurun: python tools/make_metrics_sample.py gpurun_out/metrics_sample.csv"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import test_pat_compat as T

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "metrics_sample.csv")
with tempfile.TemporaryDirectory() as d:
    text = T.build_metrics_csv(ge.load_package(), d)
open(out, "w").write(text)
print(text)
