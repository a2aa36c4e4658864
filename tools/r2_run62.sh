#!/bin/bash
for v in "ITG_X=1" "ITG_CLUSTER=1" "ITG_X=1" "ITG_CLUSTER=1"; do
env $v timeout 280 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02_lp_cfg3.json > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err || tail -3 gpurun_out/r2_bench_tmp.err
python - "$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_bench_tmp.json")); lp = json.load(open("gpurun_out/r02_lp_cfg3.json"))["launches"]
print(sys.argv[1], "cfg3", round(d["ms_per_step"], 3), "parity", d["parity"]["max_abs"], " ".join("%s=%.3f" % (x["launch"], x["ms"]) for x in lp if x["kind"] == "conv" and x["launch"].startswith(("start", "block1", "block2", "block3"))))
PY
done
