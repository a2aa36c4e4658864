#!/bin/bash
# Regenerate profiles/r02_summary.md + friends from gpurun_out/ (ncu launch list, ncu --set full raw page, bench per-launch profile, SASS histogram).
set -e
cd "$(dirname "$0")/.."
R=r02
[ -f gpurun_out/${R}_prof_cfg3.ncu-rep ] && ncu -i gpurun_out/${R}_prof_cfg3.ncu-rep --page raw --csv > gpurun_out/${R}_prof_cfg3_raw.csv 2>/dev/null
cp gpurun_out/${R}_launches_cfg3.csv profiles/${R}_launches_cfg3.csv
[ -f gpurun_out/${R}_lp_cfg3.json ] && cp gpurun_out/${R}_lp_cfg3.json profiles/${R}_launch_profile_cfg3.json
# SASS evidence: which tensor / TMA / TMEM instructions the shipped library contains, per kernel family
{
echo "# SASS opcode histogram of infinite_texture_gans_b200/libitg_b200.so (cuobjdump -sass), per kernel family"
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = TMA load, LDGSTS = cp.async, HMMA = mma.sync, SYNCS = mbarrier"
cuobjdump -sass infinite_texture_gans_b200/libitg_b200.so | python3 -c '
import sys, re, collections
fam = None; counts = collections.defaultdict(collections.Counter)
ops = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "LDGSTS", "HMMA", "SYNCS", "R2UR", "ELECT"]
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        n = m.group(1)
        fam = next((k for k in ("ssm_fused2", "ssm_fused", "conv_pair", "conv_tile", "conv_umma", "attention_mma", "conv_direct", "attention_kernel", "halo_xchg", "noise_normal", "image_to_u8") if k in n), "other")
        counts[fam]["kernels"] += 1
        continue
    if fam is None: continue
    for o in ops:
        if re.search(r"\b" + re.escape(o) + r"\b", line):
            counts[fam][o] += 1
            break
print("family".ljust(16) + "kernels".rjust(8) + "".join(o.rjust(14) for o in ops))
for f, c in sorted(counts.items()):
    print(f.ljust(16) + str(c["kernels"]).rjust(8) + "".join(str(c[o]).rjust(14) for o in ops))
'
} > profiles/${R}_sass_histogram.txt
{
echo "# Round 2 — profile summary (B200, cfg3 = 34 Generator: n_layers_G = 5, SSM, attention, 61x61 patch grid = 3904x3904, fp16 operands)"
echo
echo "Raw material: \`${R}_launches_cfg3.csv\` (ncu \`--metrics gpu__time_duration.sum --clock-control none\` over \`python bench.py --workload cfg3 --steps 2 --warmup 3 --no-graph --no-cpu-baseline\`, after the same command exited 0 without ncu), \`${R}_launch_profile_cfg3.json\` (CUDA-event time of every launch of one step, measured live by \`bench.py --profile-out\`), one \`ncu --set full --clock-control none --import-source on\` capture of the block-4/5 launches of one pass (\`tools/run_plan.py --workload cfg3\`; the 55 MB \`.ncu-rep\` stays in gpurun_out/, its raw page is tabulated below) and \`${R}_sass_histogram.txt\`. Regenerate with \`tools/make_profile_summary.sh\`."
echo
echo "## ncu launch list: share of each kernel (cold-cache, serialised; compare shares, not absolutes)"
echo
python tools/ncu_summary.py launches gpurun_out/${R}_launches_cfg3.csv
echo
echo "(\`ssm_fused2_kernel\`: StochasticSpatialModulation on CTA pairs, \`tcgen05.mma.cta_group::2\`; \`conv_pair_kernel<T, F, MODE>\`: 3x3 (MODE 0) and 1x1 (MODE 1) convs with 64 <= k_pad <= 128 on CTA pairs; \`conv_tile_kernel<T, F, MODE>\`: F = epilogue flags RES=1 RAW=2 ACT=4 IMG=8; MODE 0 = 3x3, 1 = 1x1, 2 = folded up-sampling conv; \`conv_umma_kernel<T, F>\` likewise. The FillFunctor launch is bench.py's 256 MiB L2 flush.)"
echo
echo "## CUDA-event time per launch of one step (bench.py, eager launches, L2 warm)"
echo
echo '```'
python tools/show_profile.py gpurun_out/${R}_lp_cfg3.json
echo '```'
echo
echo "## ncu --set full (16 consecutive SSM / pair / thin-layer launches of the second pass: blocks 4-5 and the final conv)"
echo
python tools/ncu_summary.py raw gpurun_out/${R}_prof_cfg3_raw.csv
} > profiles/${R}_summary.md
echo "wrote profiles/${R}_summary.md"
