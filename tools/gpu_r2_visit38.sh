#!/bin/bash
# Round-2 visit 38 (1 GPU): sigma DMMA kernel with six warps per CTA (3 row parts x 2) against four (2 x 2): parity tests, A/B
out=gpurun_out; mkdir -p $out; tag=r3r
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sigma or spin_flip or config2 or fused_energy or kernel_variants" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log | cut -c1-250
timeout 300 python tools/ab_option.py 16 sgemm_wm 2 3 tups > $out/${tag}_ab_sgemm_wm.txt 2>&1; cat $out/${tag}_ab_sgemm_wm.txt
timeout 300 python tools/ab_option.py 16 sgemm_wm 2 3 > $out/${tag}_ab_sgemm_wm_full.txt 2>&1; cat $out/${tag}_ab_sgemm_wm_full.txt
