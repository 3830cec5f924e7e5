#!/bin/bash
# usage: tools/multi.sh N shape k [extra bench args]   - one torchrun bench on N GPUs of this box, one summary line
N=$1; SHAPE=$2; K=$3; shift 3
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
  bench.py --gpus $N --steps 5 --warmup 3 --shape $SHAPE --k $K "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$SHAPE k=$K x$N | qps', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'], 3), 'k1 ms', round(d['roofline']['kernel_ms'], 3), 'build_s', d['index_build_s'], 'exchange', d.get('exchange'))" || echo "$SHAPE x$N FAILED"
