#!/usr/bin/env python
"""one-line summary of bench.py JSON lines: scripts/bshow.py label file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r = j.get("roofline", {})
        st = r.get("stage_ms_per_step") or j.get("stage_ms_per_step")
        print(f, "%.1f M rays/s" % (j["value"] / 1e6), "ms %.3f" % j["ms_per_step"], "e2e %.1f" % (j.get("e2e", {}).get("value", 0) / 1e6),
              {k: round(v, 3) for k, v in (st or {}).items()}, "err", j.get("max_abs_err_vs_oracle_512rays"))
    except Exception as e:
        print(f, "ERR", e)
