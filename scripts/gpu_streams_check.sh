#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x > gpurun_out/r02_streams_tests.log 2>&1
tail -6 gpurun_out/r02_streams_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-micro > gpurun_out/r02_streams2_bench.json 2> gpurun_out/r02_streams2_bench.err
python - <<EOF2
import json
try:
    d = json.loads(open("gpurun_out/r02_streams2_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"], d["clocks"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_streams2_bench.err").read()[-2500:])
EOF2
