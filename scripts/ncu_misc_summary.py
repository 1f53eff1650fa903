"""Per-kernel table of an `ncu --set full --page raw --csv` export: time, DRAM bytes and rate, L1 hit
rate, achieved occupancy, top stall reason -- profiles/r02_ncu_misc_summary.json."""
import csv
import json
import sys


def main(path):
    with open(path) as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    hdr, units, body = rows[0], rows[1], [r for r in rows[2:] if len(r) == len(rows[0])]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, key):
        i = col.get(key)
        if i is None or r[i] in ("", "n/a"):
            return None
        return float(r[i].replace(",", ""))

    def scaled(r, key, to):
        v = num(r, key)
        if v is None:
            return None
        u = units[col[key]]
        f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
             "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(u, 1.0)
        return v * f

    stalls = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    out = []
    for r in body:
        rd, wr = scaled(r, "dram__bytes_read.sum", 1), scaled(r, "dram__bytes_write.sum", 1)
        ms = scaled(r, "gpu__time_duration.sum", 1)
        top = None
        if stalls:
            best = max(stalls, key=lambda h: num(r, h) or 0.0)
            top = best.split("issue_stalled_")[1].split("_per_warp")[0]
        out.append({
            "kernel": r[col["Kernel Name"]][:80], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
            "ms": ms, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_GBps": (rd + wr) / ms / 1e6 if ms and rd is not None else None,
            "l1_hit_pct": num(r, "l1tex__t_sector_hit_rate.pct"),
            "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
            "achieved_occupancy_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "sm_throughput_pct": num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_throughput_pct": num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": num(r, "launch__registers_per_thread"), "top_stall": top})
    json.dump({"source": path, "launches": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
