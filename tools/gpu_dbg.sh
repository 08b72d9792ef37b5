#!/bin/bash
# timing-only diagnostics of the tensor-core forward (results are wrong on purpose with DEBUG != 0)
mkdir -p gpurun_out
for cfg in ${CFGS:-"2 0" "2 1" "2 2" "2 3" "1 0"}; do
  set -- ${cfg/_/ }
  DEEPJOIN_B200_DEBUG=$2 timeout 300 python bench.py --steps 3 --warmup 3 --query-only > gpurun_out/dbg_$1_$2.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/dbg_$1_$2.log') if x.startswith('{')][-1]; d=json.loads(l)
    print('TC=$1 DEBUG=$2 %.4g q/s grid %.2f ms  %.0f TF clocks %s' % (d['queries_per_s'], d['grid_query_ms'], d['tflops_alg'], d['clocks']))
except Exception as e:
    print('TC=$1 DEBUG=$2 failed', e); print(open('gpurun_out/dbg_$1_$2.log').read()[-800:])
PY
done
