#!/bin/bash
for f in 3 5 7; do
ITG_PAIR_INFLIGHT=$f ITG_CONV_PAIR=2 timeout 280 python bench.py --workload cfg5band --steps 10 --warmup 3 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02_lp_cfg5band.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
python - $f <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_bench_tmp.json")); lp = json.load(open("gpurun_out/r02_lp_cfg5band.json"))["launches"]
print("INFLIGHT=" + sys.argv[1], round(d["ms_per_step"], 4), " ".join("%s=%.3f" % (x["launch"], x["ms"]) for x in lp if x["kind"] == "conv" and x["launch"].startswith(("block4", "block5", "block6"))))
PY
done
