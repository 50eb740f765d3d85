#!/bin/bash
# Round-2 visit 24 (1 GPU): sigma panel kernels on per-string partner tables (etab = tab): parity, A/B at CAS(16,16), launch list.
out=gpurun_out; mkdir -p $out; tag=r2x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sigma or rdm or config2" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $out/${tag}_tests.log
timeout 600 python tools/ab_option.py 16 etab smem tab > $out/${tag}_ab_etab_tab.txt 2>&1; cat $out/${tag}_ab_etab_tab.txt
