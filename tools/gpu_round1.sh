#!/bin/bash
# First GPU pass: diagnostics per stage, then the gpu test files one by one, then a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
for t in forward mc latent mesh; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -q -m gpu -x --timeout 300 > gpurun_out/pytest_$t.log 2>&1
  echo "pytest $t rc=$?" >> gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/diag.log; tail -3 gpurun_out/bench.log
