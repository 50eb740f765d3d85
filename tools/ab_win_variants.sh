#!/bin/bash
# A/B of compile-time variants of win_kernel (libraries built in the container with -D..., loaded through SQSV_LIB): per-step time of the
# bench circuit (tools/win_scan.py) for the default library and every tools/libsqsv_*.so present.
out=gpurun_out; mkdir -p $out; tag=${1:-r3z}
{
echo "== default"; timeout 200 python tools/win_scan.py 2>&1 | tail -1
for lib in tools/libsqsv_*.so; do
  echo "== $lib"; SQSV_LIB=$PWD/$lib timeout 200 python tools/win_scan.py 2>&1 | tail -1
done
echo "== default (again)"; timeout 200 python tools/win_scan.py 2>&1 | tail -1
} > $out/${tag}_ab_win_variants.txt 2>&1
cat $out/${tag}_ab_win_variants.txt
