"""Wall time of the production run (configs[1]) with and without the latency kernel, warm process (development aid)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import run_production_tmcmc as rp

for rep in range(3):
    for dev, lat in ((False, False), (False, True), (True, False), (True, True)):
        out, s, ll = rp.run_condition("Dysbiotic_HOBIC_legacy_1k30", 1000, 2, device_mutation=dev, latency_kernel=lat)
        print(json.dumps({"rep": rep, "device_mutation": dev, "latency_kernel": lat, "wall_s": round(out["wall_s"], 3),
                          "likelihood_wall_s": round(out["likelihood_wall_s"], 3), "evals": out["likelihood_evaluations"]}), flush=True)
