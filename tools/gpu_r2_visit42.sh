#!/bin/bash
# Round-2 visit 42 (1 GPU): persistent sigma DMMA kernel (copy pipeline running over the tile boundaries): parity tests, A/B
out=gpurun_out; mkdir -p $out; tag=r3v
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sigma or spin_flip or config2 or fused_energy or kernel_variants or table_free" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log | cut -c1-250
timeout 300 python tools/ab_option.py 16 sgemm_persistent 0 1 tups > $out/${tag}_ab_sgemm_persistent.txt 2>&1; cat $out/${tag}_ab_sgemm_persistent.txt
timeout 300 python tools/ab_option.py 16 sgemm_persistent 0 1 > $out/${tag}_ab_sgemm_persistent_full.txt 2>&1; cat $out/${tag}_ab_sgemm_persistent_full.txt
