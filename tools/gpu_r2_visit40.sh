#!/bin/bash
# Round-2 visit 40 (1 GPU): sigma after the DMMA kernel skips its padded fragments (A/B line of the default), setter timing in the bench
out=gpurun_out; mkdir -p $out; tag=r3t
timeout 300 python tools/ab_option.py 16 sgemm_wm 2 3 tups > $out/${tag}_ab_sgemm_wm.txt 2>&1; cat $out/${tag}_ab_sgemm_wm.txt
timeout 600 python bench.py --no-extras --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3t_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["batched"]["value"], d["e2e"]["wavefunction_setter"])
PY
