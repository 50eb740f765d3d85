#!/bin/bash
# Round-2 seventh visit (1 GPU): symmetric S / A Gram route of the 2-RDM -- parity, A/B timing; bench sanity after the batch template.
out=gpurun_out; mkdir -p $out; tag=r2g
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_saups.py tests/test_gpu_linear_response.py -m gpu -x -q > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -6 $out/${tag}_tests.log
timeout 300 python tools/ab_option.py 16 rdm_sym 0 1 > $out/${tag}_ab_rdm_sym.txt 2>&1; tail -10 $out/${tag}_ab_rdm_sym.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 400 $out/${tag}_bench.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_sym_kernel -s 4 -c 1 -f -o $out/${tag}_gram_sym python tools/ab_option.py 16 rdm_sym 1 1 > /dev/null 2>&1
ls -la $out/${tag}_gram_sym.ncu-rep
