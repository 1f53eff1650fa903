"""Join an `ncu --set full` capture of the tensor-core kernels of one bench step with the launch
manifest bench.py wrote (`--dump-igemm`), and emit the per-launch table + means that
`bench.py` reports as `roofline.traffic` (profiles/r02_ncu_tc_summary.json).

  ncu --set full --clock-control none --import-source on \\
      -k regex:"igemm_tc_kernel|conv_halo_tc_kernel" --nvtx --nvtx-include "timed/" -c 28 \\
      -o gpurun_out/r02_ncu_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline \\
      --dump-igemm gpurun_out/igemm_manifest.json
  ncu -i gpurun_out/r02_ncu_tc.ncu-rep --page raw --csv > profiles/r02_ncu_tc_raw.csv
  python scripts/ncu_tc_summary.py profiles/r02_ncu_tc_raw.csv gpurun_out/igemm_manifest.json \\
      > profiles/r02_ncu_tc_summary.json
"""
import csv
import json
import sys

LAYERS = ["conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2",
          "conv4_3", "conv5_1", "conv5_2", "conv5_3", "rpn_conv_3x3", "rpn_cls_score|rpn_bbox_pred"]
HEAD = ["fc6_maskest", "mask_pred", "fc6", "fc6_mask", "fc7", "fc7_mask", "cls_score|seg_cls_score|bbox_pred"]


def main(raw_csv, manifest_json):
    with open(raw_csv) as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    with open(manifest_json) as f:
        man = json.load(f)
    per_step = len(man["launches"]) // man["steps"]
    launches = man["launches"][:per_step]
    names = LAYERS + [h + s for s in ("", "_ext") for h in HEAD]
    assert len(body) >= per_step, "capture holds %d launches, one step has %d" % (len(body), per_step)

    def val(r, key, scale=1.0):
        return float(r[col[key]].replace(",", "")) * scale if key in col else None

    def unit(key):
        return rows[1][col[key]] if key in col else ""

    def to_bytes(r, key):
        u = unit(key).lower()
        mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return val(r, key) * mul

    def to_ms(r, key):
        u = unit(key).lower()
        mul = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        return val(r, key) * mul

    out = []
    for i, (r, m) in enumerate(zip(body[:per_step], launches)):
        ms = to_ms(r, "gpu__time_duration.sum")
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        out.append({
            "layer": names[i] if i < len(names) else "launch%d" % i,
            "kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mnc::", ""),
            "gemm_MxNxK": "%dx%dx%d" % (m["M"], m["N"], m["K"]), "split_k": m["split_k"],
            "pooled_epilogue": m["pooled"],
            "duration_ms": ms,
            "tensor_pipe_active_pct": val(r, "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active")
            or val(r, "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active")
            or val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
            "algorithmic_bytes": m["bytes"], "algorithmic_flops": m["flops"],
            "algorithmic_TFLOPs": m["flops"] / (ms * 1e-3) / 1e12,
            "registers": val(r, "launch__registers_per_thread"),
        })
    n = len(out)
    summary = {
        "source": "ncu --set full --clock-control none over the tensor-core launches of one bench.py step "
                  "(batch 8, 600x1000); raw: profiles/r02_ncu_tc_raw.csv; made by scripts/ncu_tc_summary.py",
        "n_launches": n,
        "mean_traffic_bytes_per_launch": sum(o["traffic_bytes"] for o in out) / n,
        "mean_algorithmic_bytes_per_launch": sum(o["algorithmic_bytes"] for o in out) / n,
        "sum_duration_ms": sum(o["duration_ms"] for o in out),
        "sum_algorithmic_flops": sum(o["algorithmic_flops"] for o in out),
        "launches": out,
    }
    json.dump(summary, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
