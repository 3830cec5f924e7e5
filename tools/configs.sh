#!/bin/bash
# The BASELINE.json configurations that fit one GPU, one bench.py run each (short): prints one summary line per run.
run() {
  timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 3 "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', '| qps', round(d['value']), 'e2e', round(d['e2e']['value']), 'py', round(d.get('e2e_python', {}).get('value', 0)), 'ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'n_general', d['n_general'])" || echo "$* FAILED"
}
run --shape msmarco --k 10
run --shape msmarco --k 100
run --shape msmarco --k 1000
run --shape nq --k 10
run --shape nq --k 10 --method bm25+ --queries 10000
run --shape nq --k 10 --method bm25l --queries 10000
run --shape scifact --k 10
