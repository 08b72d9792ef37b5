#!/bin/bash
# quick iteration: UMMA probe, correctness of the tensor-core path, then bench per cluster size
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py probe > gpurun_out/diag_probe.log 2>&1; tail -c 1500 gpurun_out/diag_probe.log; echo
for C in ${CLUSTERS:-1 2}; do
  DEEPJOIN_B200_CLUSTER=$C timeout 300 python tools/gpu_diag.py tc > gpurun_out/diag_tc_c$C.log 2>&1
  DEEPJOIN_B200_CLUSTER=$C timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$C.log 2>&1
  echo "C=$C"; tail -c 900 gpurun_out/diag_tc_c$C.log | head -c 600; echo; python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/bench_c$C.log') if x.startswith('{')][-1]; d=json.loads(l)
    print('bench C=$C value %.4g q/s  grid %.2f ms mc %.2f ms  frac %.3f  e2e %.4g clocks %s' % (d['value'], d['config']['grid_query_ms'], d['config']['marching_cubes_ms'], d['roofline']['frac'], d['e2e']['value'], d['clocks']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_c$C.log').read()[-1500:])
PY
done
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_mesh.py -q -m gpu -x --timeout 300 > gpurun_out/pytest_fwd_mesh.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_fwd_mesh.log
fi
