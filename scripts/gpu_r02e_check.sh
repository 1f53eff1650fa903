#!/bin/bash
# Round-2 (session 3, third pass): the NMS tests across all four forms, timings of the forms,
# voting launch-shape A/B, 14x14 row-walk planes A/B, ncu of the wide cluster NMS.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nms.py tests/test_gpu_voting.py tests/test_gpu_proposal.py tests/test_gpu_roi.py tests/test_ref_pin.py -m gpu -q > gpurun_out/r02e_pytest_subset.log 2>&1
tail -5 gpurun_out/r02e_pytest_subset.log
timeout 200 python scripts/gpu_nms_modes.py > gpurun_out/r02e_nms_modes.log 2>&1; tail -3 gpurun_out/r02e_nms_modes.log
timeout 300 python scripts/gpu_mv_shape_ab.py > gpurun_out/r02e_mv_shape_ab.log 2>&1; tail -22 gpurun_out/r02e_mv_shape_ab.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_" --csv --log-file gpurun_out/r02e_nms_modes_launches.csv python scripts/gpu_nms_modes.py > /dev/null 2>&1
python - <<EOF2
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/r02e_nms_modes_launches.csv") if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); g = hdr.index("Grid Size")
agg = {}
for r in rows[1:]:
    key = (r[k][:40], r[g])
    agg.setdefault(key, []).append(float(r[v].replace(",", "")))
for key, vals in agg.items():
    vals.sort()
    print(key, "n=%d median %.1f us" % (len(vals), vals[len(vals) // 2] / 1000.0))
EOF2
