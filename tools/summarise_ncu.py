"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep itself stays in gpurun_out/)."""
import collections, csv, subprocess, sys

def launches(csv_path, out_path, note=""):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out_path, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list, summed per kernel. {note}\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':50s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share_%':>8s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:50s} {n:8d} {t:12.1f} {t / n:10.1f} {t / tot * 100:8.1f}\n")
        f.write(f"{'TOTAL':50s} {'':8s} {tot:12.1f}\n")

def full(rep_path, out_path, note=""):
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "smsp__warps_eligible.avg.per_cycle_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
    with open(out_path, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch per row. {note}\n")
        for r in rows[2:]:
            d = dict(zip(h, r))
            for k in want:
                if k in d:
                    f.write(f"{k:70s} {d[k]:>20s} {units[h.index(k)]}\n")
            st = {k: float(v.replace(",", "")) for k, v in d.items() if "pcsamp_warps_issue_stalled" in k and not k.endswith("_not_issued") and v}
            tot = sum(st.values()) or 1
            f.write("stall reasons (pc samples):\n")
            for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
                f.write(f"   {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):40s} {v / tot * 100:6.1f} %\n")
            f.write("\n")

if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    (launches if kind == "launches" else full)(src, dst, note)
