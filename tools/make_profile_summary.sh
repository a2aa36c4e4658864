#!/bin/bash
# Regenerate profiles/r01_summary.md + friends from gpurun_out/ (ncu launch list, ncu --set full raw page, bench profile).
set -e
cd "$(dirname "$0")/.."
ncu -i gpurun_out/r01_prof_cfg2.ncu-rep --page raw --csv > gpurun_out/r01_prof_cfg2_raw.csv 2>/dev/null
cp gpurun_out/r01_launches.csv profiles/r01_launches_cfg2.csv
[ -f gpurun_out/launch_profile_cfg2.json ] && cp gpurun_out/launch_profile_cfg2.json profiles/r01_launch_profile_cfg2.json
{
echo "# Round 1 — profile summary (B200, cfg2 = 241 Generator, 7x21 patch grid, fp16 operands)"
echo
echo "Raw material: \`r01_launches_cfg2.csv\` (ncu \`--metrics gpu__time_duration.sum --clock-control none\` over \`python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline\`, after the same command exited 0 without ncu), \`r01_launch_profile_cfg2.json\` (CUDA-event time of every launch of one step, measured live by \`bench.py --profile-out\`), and one \`ncu --set full --clock-control none --import-source on\` capture of one whole pass (\`tools/run_plan.py\`; the 47 MB \`.ncu-rep\` stays in gpurun_out/, its raw page is tabulated below). Regenerate with \`tools/make_profile_summary.sh\`."
echo
echo "## ncu launch list: share of each kernel (cold-cache, serialised; compare shares, not absolutes)"
echo
python tools/ncu_summary.py launches gpurun_out/r01_launches.csv
echo
echo "(\`conv_tile_kernel<T, F, MODE>\`: F = epilogue flags RES=1 RAW=2 ACT=4 IMG=8; MODE 0 = 3x3, 1 = 1x1, 2 = folded up-sampling conv; \`conv_umma_kernel<T, F>\` likewise. The FillFunctor launch is bench.py's 256 MiB L2 flush.)"
echo
echo "## CUDA-event time per launch of one step (bench.py, eager launches, L2 warm)"
echo
echo '```'
python tools/show_profile.py gpurun_out/launch_profile_cfg2.json
echo '```'
echo
echo "## ncu --set full, one pass (ids in launch order: start, block1.conv1/2, block2.conv3/1/2, block3.conv3/1/2, attention, block4.conv3/1, then the halo-tile launches block4.conv2, block5.conv3/1/2, block6.conv3/1/2, final)"
echo
python tools/ncu_summary.py raw gpurun_out/r01_prof_cfg2_raw.csv
} > profiles/r01_summary.md
python - <<'PY'
import csv, json
rows = list(csv.reader(open('gpurun_out/r01_prof_cfg2_raw.csv')))
hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
mult = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9}
conv = 0
for r in rows[2:]:
    b = float(r[idx['dram__bytes_read.sum']]) * mult[units[idx['dram__bytes_read.sum']]] + float(r[idx['dram__bytes_write.sum']]) * mult[units[idx['dram__bytes_write.sum']]]
    if 'attention' not in r[idx['Kernel Name']]:
        conv += b
d = json.load(open('profiles/r01_traffic.json'))
d['cfg2:fp16'] = int(conv)
json.dump(d, open('profiles/r01_traffic.json', 'w'), indent=1)
print('conv DRAM traffic per pass:', conv / 1e6, 'MB')
PY
