#!/usr/bin/env python
import json, sys
j = json.load(open(sys.argv[1]))
print("value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "surfels", j["config"]["surfels_after_step"],
      "ms/step", round(j["ms_per_step"], 2), j["clocks"], "launches", j.get("gpu_launches"))
for k, v in j.get("kernels", {}).items():
    print(f"{k:24s} {v['mean_us']:8.2f} us  {v['share'] * 100:5.1f}%")
if "roofline" in j:
    r = j["roofline"]
    print("roofline", r["kernel"], round(r["achieved"], 1), "GB/s frac", round(r["frac"], 4))
if "cpu_baseline" in j:
    print("cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"])
