#!/bin/bash
# usage: tools/sweep.sh "THREADS CTAS STAGE" ...   (runs bench.py once per configuration, prints ms_per_step)
for cfg in "$@"; do
  set -- $cfg
  BM25CU_THREADS=$1 BM25CU_CTAS_PER_SM=$2 BM25CU_POOL=$3 BM25CU_CBMIN=${4:-512} timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg', '$cfg', 'ms', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))" || echo "cfg $cfg FAILED"
done
