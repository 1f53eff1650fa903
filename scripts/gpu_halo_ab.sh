#!/bin/bash
# A/B of the halo kernel's two precision modes on the GPU box (run under gpurun from the repo root).
timeout 60 scripts/exp/umma_offset_sw64_fp8_test > gpurun_out/r02_umma_sw64_fp8_shifted_window_experiment.log 2>&1; tail -3 gpurun_out/r02_umma_sw64_fp8_shifted_window_experiment.log
timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_e2e.py tests/test_gpu_round2.py -m gpu -q -x > gpurun_out/r02_halo_pm1_tests.log 2>&1
tail -15 gpurun_out/r02_halo_pm1_tests.log
timeout 200 python bench.py --no-cpu-baseline --no-micro > gpurun_out/r02_halo_pm1_bench.json 2> gpurun_out/r02_halo_pm1_bench.err
timeout 200 python bench.py --no-cpu-baseline --no-micro --halo-split > gpurun_out/r02_halo_split_bench.json 2>/dev/null
python - <<EOF2
import json
for f in ["r02_halo_pm1_bench", "r02_halo_split_bench"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["tensor_work_factor"])
        print(d["roofline"]["ms_by_launch_site"][:6])
    except Exception as e:
        print(f, "ERR", e)
        print(open("gpurun_out/r02_halo_pm1_bench.err").read()[-2000:])
EOF2
