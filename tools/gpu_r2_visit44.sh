#!/bin/bash
# Round-2 visit 44 (1 GPU): light-cone state construction from a reference determinant: parity tests, wave-function goldens, setter timing
out=gpurun_out; mkdir -p $out; tag=r3y
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "light_cone or wavefunction or config2 or rotosolve or tups_energy" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log | cut -c1-250
timeout 600 python bench.py --no-extras --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3y_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["batched"]["value"], json.dumps(d["e2e"]["wavefunction_setter"])[:900])
PY
