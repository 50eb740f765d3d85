#!/bin/bash
# Round-2 visit 37 (1 GPU): sigma of a tUPS state on the blocked panels: DMMA residency 2 vs 1 CTAs per SM, pipeline on / off
out=gpurun_out; mkdir -p $out; tag=r3q
timeout 300 python tools/ab_option.py 16 sgemm_cta 2 1 tups > $out/${tag}_ab_sgemm_cta.txt 2>&1; cat $out/${tag}_ab_sgemm_cta.txt
timeout 300 python tools/ab_option.py 16 pipeline 1 0 tups > $out/${tag}_ab_pipeline.txt 2>&1; cat $out/${tag}_ab_pipeline.txt
