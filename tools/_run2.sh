for r in 8 11 16; do echo "range $r"; SQ_WIN_RANGE=$r timeout 200 python tools/win_scan.py "1" 2>&1 | tail -1; done
