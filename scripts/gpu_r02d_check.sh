#!/bin/bash
# Round-2 (session 3, second pass) on the box: every -m gpu test, launch-shape A/B of the row-walk
# ROIWarping kernel, the default bench line, the bench with the single-CTA capped NMS (A/B), and an
# ncu --set full capture of the kernels that changed (capped NMS, voting, softmax, row walk).
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02d_pytest.log 2>&1
tail -12 gpurun_out/r02d_pytest.log
timeout 200 python scripts/gpu_roi_walk_shape_ab.py > gpurun_out/r02d_roi_walk_shape_ab.log 2>&1
cat gpurun_out/r02d_roi_walk_shape_ab.log | tail -24
timeout 400 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
timeout 300 python bench.py --no-micro --no-cpu-baseline --nms-single-cta > gpurun_out/r02d_bench_ab_nms_single_cta.json 2> gpurun_out/r02d_bench_ab.err
python - <<EOF2
import json
for f in ("gpurun_out/r02d_bench.json", "gpurun_out/r02d_bench_ab_nms_single_cta.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["value"], d["ms_per_step"], r["frac"], d["e2e"]["value"], d["latency_batch1_ms"],
              d["forward_plus_voting"]["value"], d["gpu_launches_per_step"], d["clocks"])
        if d.get("micro"):
            print({k: (v.get("ms"), v.get("frac_of_hbm")) for k, v in d["micro"].items()})
    except Exception as e:
        print("ERR", f, e)
EOF2
tail -5 gpurun_out/r02d_bench.err gpurun_out/r02d_bench_ab.err
timeout 500 ncu --set full --clock-control none --import-source on \
  --nvtx --nvtx-include "profile/" -k regex:"nms_lazy|mv_aggregate|vote_select|softmax_rows|roi_warp_rowwalk|mv_finalize" -c 24 \
  -o /tmp/r02d python scripts/gpu_profile_misc.py > gpurun_out/r02d_ncu.log 2>&1
tail -3 gpurun_out/r02d_ncu.log
ncu -i /tmp/r02d.ncu-rep --page raw --csv > gpurun_out/r02d_ncu_raw.csv 2>/dev/null
python scripts/ncu_misc_summary.py gpurun_out/r02d_ncu_raw.csv > gpurun_out/r02d_ncu_summary.json 2>gpurun_out/r02d_ncu_summary.err
python - <<EOF3
import json
try:
    d = json.load(open("gpurun_out/r02d_ncu_summary.json"))
    for r in d["launches"]:
        print("%-60s %-14s %8.3f ms occ %5.1f sm %5.1f l1 %5.1f" % (r["kernel"][:60], r["grid"], r["ms"], r["achieved_occupancy_pct"] or 0, r["sm_throughput_pct"] or 0, r["l1tex_throughput_pct"] or 0))
except Exception as e:
    print("ERR", e)
EOF3
