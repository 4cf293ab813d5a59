#!/usr/bin/env python
"""Derive cra5_b200/api/era5_268v_stats.json (one record per model channel: name, mean, std) from the ERA5
climatology tables shipped with the reference (cra5/api/mean_std.json, mean_std_single.json), applying the channel
order and level selection of cra5_api.get_mean_std (cra5_api.py:33-34, 243-261). Build-container only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CRA5_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
from cra5_b200.api import era5_268v as V  # noqa: E402

with open(os.path.join(REF, "cra5/api/mean_std.json")) as f:
    pl = json.load(f)
with open(os.path.join(REF, "cra5/api/mean_std_single.json")) as f:
    sl = json.load(f)

records = []
for v in V.PRESSURE_VARS:
    for li, level in enumerate(V.PRESSURE_LEVELS):  # the stats tables are indexed by position in the 37-level list
        records.append({"name": f"{v}_{int(level)}", "mean": pl["mean"][v][li], "std": pl["std"][v][li]})
for v in V.SINGLE_VARS:
    records.append({"name": v, "mean": sl["mean"][v], "std": sl["std"][v]})
assert len(records) == 268
out = os.path.join(ROOT, "cra5_b200", "api", "era5_268v_stats.json")
with open(out, "w") as f:
    json.dump({"source": "ERA5 per-channel climatology (CRA5 release)", "channels": records}, f, indent=0)
print("wrote", out)
