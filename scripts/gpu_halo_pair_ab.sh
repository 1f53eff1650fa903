#!/bin/bash
# Halo-kernel CTA pairs: parity tests, then A/B bench (run under gpurun from the repo root).
timeout 240 python -m pytest tests/test_gpu_round2.py tests/test_gpu_dense.py -m gpu -q -x > gpurun_out/r02_halo_pair_tests.log 2>&1
tail -8 gpurun_out/r02_halo_pair_tests.log
timeout 200 python bench.py --no-cpu-baseline --no-micro > gpurun_out/r02_halo_pair_bench.json 2> gpurun_out/r02_halo_pair_bench.err
timeout 200 python bench.py --no-cpu-baseline --no-micro --halo-single > gpurun_out/r02_halo_single_bench.json 2>/dev/null
python - <<EOF2
import json
for f in ["r02_halo_pair_bench", "r02_halo_single_bench"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["tensor_work_factor"])
        print(d["roofline"]["ms_by_launch_site"][:4])
    except Exception as e:
        print(f, "ERR", e)
        print(open("gpurun_out/r02_halo_pair_bench.err").read()[-2000:])
EOF2
