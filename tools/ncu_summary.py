"""Summarise ncu outputs into profiles/: (a) a launch list CSV -> per-kernel totals/shares, (b) a --set full report ->
the roofline-relevant metrics per captured launch."""
import collections, csv, re, subprocess, sys

def launches(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        k = re.sub(r"\(.*", "", row["Kernel Name"]) + " grid=" + row.get("Grid Size", "") + " block=" + row.get("Block Size", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as o:
        o.write(f"# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n")
        o.write("kernel | launches | total_us | avg_us | share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"{k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f}%\n")
        o.write(f"TOTAL | {sum(v[0] for v in agg.values())} | {tot:.1f} | |\n")

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]

def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w") as o:
        o.write(f"# source: {path} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            o.write("----\n")
            for i in idx:
                o.write(f"{hdr[i]} = {r[i]} {units[i]}\n")

if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (launches if kind == "launches" else full)(src, dst)
