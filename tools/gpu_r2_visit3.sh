#!/bin/bash
# Round-2 third visit (1 GPU): hand-written DMMA kernels in place of cuBLAS -- parity suite, energy timings.
out=gpurun_out; mkdir -p $out; tag=r2c
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -12 $out/${tag}_tests.log
timeout 300 python tools/bench_energy.py 16 2 > $out/${tag}_energy.txt 2>&1; tail -12 $out/${tag}_energy.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_energy_launches.csv python tools/bench_energy.py 16 1 > $out/${tag}_energy_ncu.log 2>&1
python - <<'P'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2c_energy_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    a=agg[r[ki][:60]]; a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]:
    print(f"{k:60s} n={a[0]:4d} total={a[1]/1e6:9.2f} ms share={a[1]/tot:.3f} avg={a[1]/a[0]/1e3:.1f} us")
P
