#!/bin/bash
# Round-2 visit 33 (1 GPU): blocked panels (32 x 32 blocks of determinants) of the half sigma / RDM build: parity, A/B at CAS(16,16),
# launch list of one sigma build.
out=gpurun_out; mkdir -p $out; tag=r3k
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spin_flip or sigma or rdm or config2 or fused_energy or backwards or wavefunction" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -12 $out/${tag}_tests.log | cut -c1-250
timeout 600 python tools/ab_sigma_spinsym.py > $out/${tag}_ab_spinsym.txt 2>&1; cat $out/${tag}_ab_spinsym.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_sigma_launches.csv python tools/ab_sigma_spinsym.py 14 > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r3k_sigma_launches.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn = H.index("Kernel Name"); mv = H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    a = agg[r[kn][:60]]; a[0] += 1; a[1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{k:60s} n={c:4d} total={t/1e6:9.3f} ms avg={t/c/1e3:9.1f} us")
PY
