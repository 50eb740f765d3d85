#!/bin/bash
# Round-2 fifth visit (1 GPU): thin-CTA sigma DMMA kernel -- parity, A/B of one vs two GEMM CTAs per SM, pipeline on/off.
out=gpurun_out; mkdir -p $out; tag=r2e
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sigma or rdm or energy or config2 or config3 or gradient" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $out/${tag}_tests.log
timeout 300 python tools/ab_option.py 16 sgemm_cta 1 2 > $out/${tag}_ab_sgemm_cta.txt 2>&1; tail -10 $out/${tag}_ab_sgemm_cta.txt
timeout 300 python tools/ab_option.py 16 pipeline 0 1 > $out/${tag}_ab_pipeline.txt 2>&1; tail -10 $out/${tag}_ab_pipeline.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_sigma_launches.csv python tools/ab_option.py 16 pipeline 0 0 > /dev/null 2>&1
grep -E "sigma_dmma|scatter_E|build_Dsym" $out/${tag}_sigma_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -12
