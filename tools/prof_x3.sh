#!/bin/bash
# in-kernel cycle counters of the split-operand forward (debug build; tools/gpu_diag.py prof)
export DEEPJOIN_B200_LIB=$PWD/deepjoin_b200/lib/libdeepjoin_b200_debug.so
mkdir -p gpurun_out
PROF_PREC=${PROF_PREC:-fp16x2} timeout 200 python tools/gpu_diag.py --run prof > gpurun_out/prof_x3.log 2>&1
python - <<'PY'
import json
t = open("gpurun_out/prof_x3.log").read()
if "DIAG_RESULT " not in t:
    print(t[-2000:])
else:
    d = json.loads(t.split("DIAG_RESULT ")[1])
    print("ms", d.get("ms"), d.get("exception"))
    for k, v in d.items():
        if isinstance(v, dict):
            print("%-24s leader %12.0f peer %12.0f" % (k, v["leader_mean"], v["peer_mean"]))
PY
