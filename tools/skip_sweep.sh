#!/bin/bash
# timing experiment: disable parts of the fused kernel (results are wrong), print ms per batch
for m in "$@"; do
  BM25CU_DEBUG_SKIP=$m timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip', '$m', 'ms', round(d['ms_per_step'],2))" || echo "skip $m FAILED"
done
