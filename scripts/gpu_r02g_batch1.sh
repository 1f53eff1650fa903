#!/bin/bash
# Where a single image's 2.9-3.2 ms go: launch list of eager batch-1 steps + the batch-1 bench line.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 400 --csv --log-file gpurun_out/r02g_launches_batch1.csv python bench.py --batch 1 --steps 2 --warmup 3 --no-cpu-baseline --no-micro --no-graph > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02g_launches_batch1.csv > gpurun_out/r02g_launches_batch1_summary.txt 2>&1; head -40 gpurun_out/r02g_launches_batch1_summary.txt
timeout 200 python bench.py --batch 1 --streams 1 --no-cpu-baseline --no-micro > gpurun_out/r02g_bench_batch1_streams1.json 2> gpurun_out/r02g_b1.err
timeout 200 python bench.py --batch 1 --streams 1 --no-cpu-baseline --no-micro --nms-mode 0 > gpurun_out/r02g_bench_batch1_streams1_nms0.json 2>> gpurun_out/r02g_b1.err
python - <<EOF2
import json
for f in ("gpurun_out/r02g_bench_batch1_streams1.json", "gpurun_out/r02g_bench_batch1_streams1_nms0.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["latency_batch1_ms"], d["e2e"]["value"], d["clocks"])
        print([(s, m) for s, m in d["roofline"]["ms_by_launch_site"]])
    except Exception as e:
        print("ERR", f, e)
EOF2
tail -3 gpurun_out/r02g_b1.err
