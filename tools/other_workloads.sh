#!/bin/bash
# The other BASELINE.json configurations on one GPU (not the headline bench line): one JSON line each into gpurun_out/.
for w in cfg1 cfg5band cfg4 cfg3; do
  timeout 280 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_{w}.json"))
    print(w, "ms/step", round(d["ms_per_step"], 3), "MP/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv frac", round(d["roofline"]["frac"], 3),
          "step TF", round(d["roofline"]["step_tflops"]), "step frac", round(d["roofline"]["step_frac_of_peak"], 3))
except Exception as e:
    print(w, "failed", e)
PY
done
