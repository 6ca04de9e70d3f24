#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "x x" "0 x" "x 0" "0 0" "1 1"; do
  set -- $cfg
  echo "== occ3=$1 chains=$2 (x = per-code default)"
  env $( [ $1 != x ] && echo DVBS2FEC_LDPC_OCC3=$1 ) $( [ $2 != x ] && echo DVBS2FEC_LDPC_CHAINS=$2 ) timeout 600 python tools/modcod_sweep.py --out gpurun_out/tmp_sweep.json 2>&1 | python -c "
import sys,ast
o=[]
for l in sys.stdin:
    if l.startswith('{') and 'code' in l:
        r=ast.literal_eval(l); o.append('%s %.2f' % (r['code'], r['ms']))
print(' | '.join(o))"
done
