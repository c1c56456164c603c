import sys, json, time
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import numpy as np
import run_production_tmcmc as rp
from tmcmc202601_b200 import workloads
res = {}
for rep in range(2):
    for merge in (True, False):
        for dm in (True, False):
            s, out = rp.run_merged(workloads.CONDITION_KEYS, 1000, 2, device_mutation=dm, merge_launches=merge)
            res[f"rep{rep}_merge{int(merge)}_dm{int(dm)}"] = {k: s[k] for k in ("wall_s", "merged_launches", "likelihood_evaluations", "likelihood_wall_s")}
            print(rep, merge, dm, res[f"rep{rep}_merge{int(merge)}_dm{int(dm)}"], [p["stages"] for p in s["per_condition"]], flush=True)
# two chains of configs[1]
for merge in (True, False):
    s, out = rp.run_merged(["Dysbiotic_HOBIC"], 1000, 2, device_mutation=True, merge_launches=merge)
    print("configs1-like", merge, s["wall_s"], s["merged_launches"], flush=True)
json.dump(res, open("gpurun_out/exp1_streams.json", "w"), indent=1)
