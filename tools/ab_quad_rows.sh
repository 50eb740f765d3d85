#!/bin/bash
# A/B of a compile-time parameter of the gradient kernels: QUAD_ROWS (row groups per CTA of quad_grad_kernel) = 16 (default library), 32, 64.
# The variant libraries are built in the container:
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -DQUAD_ROWS=32 -I include -I slowquant_b200/csrc \
#        -o tools/libsqsv_qr32.so slowquant_b200/csrc/*.cu -lcudart -ldl
# and picked up through SQSV_LIB (slowquant_b200/_lib.py).  "x" as the only window configuration: ab_grad.py then times the one-brick and the
# two-brick sweep only.
out=gpurun_out; mkdir -p $out; tag=${1:-r3w}
{
echo "== QUAD_ROWS=16 (default)"; timeout 200 python tools/ab_grad.py 16 16 none 2>&1 | grep -E "brick|diff" | head -3
for q in ${QR_LIST:-8}; do
  echo "== QUAD_ROWS=$q"; SQSV_LIB=$PWD/tools/libsqsv_qr$q.so timeout 200 python tools/ab_grad.py 16 16 none 2>&1 | grep -E "brick|diff" | head -3
done
} > $out/${tag}_ab_quad_rows.txt 2>&1
cat $out/${tag}_ab_quad_rows.txt
