#!/bin/bash
# pair-kernel iteration: correctness (diag tc), in-kernel counters, bench
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py tc > gpurun_out/diag_tc.log 2>&1; tail -c 700 gpurun_out/diag_tc.log | head -c 400; echo
DEEPJOIN_B200_DEBUG=8 timeout 200 python tools/gpu_diag.py prof > gpurun_out/diag_prof.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/diag_prof.log') if x.startswith('prof ')][-1]
d=json.loads(l[5:])
print('prof ms %.2f' % d['ms'])
for k,v in d.items():
    if isinstance(v,dict): print('  %-22s leader %8.2fM peer %8.2fM' % (k, v['leader_mean']/1e6, v['peer_mean']/1e6))
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc2.log 2>&1
python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/bench_tc2.log') if x.startswith('{')][-1]; d=json.loads(l)
    print('bench value %.4g q/s  grid %.2f ms mc %.2f ms  frac %.3f  e2e %.4g clocks %s' % (d['value'], d['config']['grid_query_ms'], d['config']['marching_cubes_ms'], d['roofline']['frac'], d['e2e']['value'], d['clocks']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_tc2.log').read()[-1500:])
PY
if [ -n "$TESTS" ]; then
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
fi
