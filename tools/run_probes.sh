#!/bin/bash
# Build and run the micro-probes on the GPU box (through gpurun); writes gpurun_out/profiles/${PSA_ROUND:-r02}_probes.json,
# which bench.py reads for the measured integer-pipe rates (copy it to profiles/ to have it judged).
set -u
out=gpurun_out/profiles; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/alu_rate_probe tools/probes/alu_rate_probe.cu || exit 1
/tmp/alu_rate_probe | tee /tmp/alu_rate.txt
python - <<'PY'
import json, os, re
rates = {}
for line in open("/tmp/alu_rate.txt"):
    m = re.match(r"(\w+)\s+([\d.]+) lane-ops/clk/SM \((\d+) SMs", line)
    if m:
        rates[m.group(1)] = float(m.group(2))
        sms = int(m.group(3))
rec = {"lane_ops_per_clk_per_sm": rates, "sms": sms, "pipe": {"LOP3": "alu", "IADD3": "alu", "SHF": "alu", "IMAD": "fma"},
       "source": "tools/probes/alu_rate_probe.cu (one block of 32 warps per SM, 8 independent chains per thread)"}
path = os.path.join("gpurun_out/profiles", os.environ.get("PSA_ROUND", "r02") + "_probes.json")
json.dump(rec, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(rec))
PY
