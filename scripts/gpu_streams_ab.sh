#!/bin/bash
# A/B: one vs two steps in flight (run under gpurun from the repo root).
timeout 250 python bench.py --no-cpu-baseline --no-micro --streams 2 > gpurun_out/r02_streams2_bench.json 2> gpurun_out/r02_streams2_bench.err
timeout 250 python bench.py --no-cpu-baseline --no-micro --streams 1 > gpurun_out/r02_streams1_bench.json 2> gpurun_out/r02_streams1_bench.err
python - <<EOF2
import json
for f in ["r02_streams2_bench", "r02_streams1_bench"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["step_ms"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
        print(open("gpurun_out/%s.err" % f).read()[-2500:])
EOF2
