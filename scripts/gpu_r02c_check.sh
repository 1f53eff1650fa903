#!/bin/bash
# Round-2 (session 3) validation on the box: every -m gpu test, the default bench line, the same
# bench with the previous NMS / voting forms (A/B), and the launch list of two eager steps.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02c_pytest.log 2>&1
tail -14 gpurun_out/r02c_pytest.log
timeout 400 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
timeout 300 python bench.py --no-micro --no-cpu-baseline --nms-matrix --mv-full-sweep > gpurun_out/r02c_bench_ab_matrix_fullsweep.json 2> gpurun_out/r02c_bench_ab.err
python - <<EOF2
import json
for f in ("gpurun_out/r02c_bench.json", "gpurun_out/r02c_bench_ab_matrix_fullsweep.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["value"], d["ms_per_step"], r["frac"], d["e2e"]["value"], d["latency_batch1_ms"],
              d["forward_plus_voting"]["value"], d["gpu_launches_per_step"], d["clocks"])
        if d.get("micro"):
            print({k: v.get("ms") for k, v in d["micro"].items()}, d["micro"]["nms_10k_keep300"], d["micro"]["mask_voting_600x21"])
    except Exception as e:
        print("ERR", f, e)
        print(open(f.replace(".json", ".err").replace("_ab_matrix_fullsweep", "_ab")).read()[-3000:])
EOF2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --no-graph > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02c_launches.csv > gpurun_out/r02c_launches_summary.txt 2>&1; head -32 gpurun_out/r02c_launches_summary.txt
