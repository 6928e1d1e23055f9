"""Per-kernel summary of an ncu --set full report: duration, DRAM bytes (read + write) and throughput, L2 / tensor-pipe
utilisation, registers, achieved occupancy.  Usage: python scripts/ncu_summary.py gpurun_out/r2_prof_all.ncu-rep > profiles/..."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = {
    "gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
}
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
name_i = idx.get("Kernel Name")
def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return float("nan")
def to_bytes(v, unit):
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit.lower(), 1)
print(f"{'kernel':58s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>7s} {'dram%':>6s} {'L2%':>6s} {'tensor%':>7s} {'SM%':>6s} {'occ%':>6s} {'regs':>5s} {'grid':>6s}x{'blk':<4s}")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    g = {}
    for m, k in want.items():
        if m in idx:
            v = num(r[idx[m]])
            if k in ("dram_rd", "dram_wr"):
                v = to_bytes(v, units[idx[m]])
            if k == "time_us":
                v = to_us(v, units[idx[m]])
            g[k] = v
    nm = r[name_i].replace("<unnamed>::", "").replace("void ", "")[:58]
    mb = (g.get("dram_rd", 0) + g.get("dram_wr", 0)) / 1e6
    t = g.get("time_us", float("nan"))
    print(f"{nm:58s} {t:9.1f} {mb:9.1f} {mb / t * 1e3 if t else 0:7.0f} {g.get('dram_pct', float('nan')):6.1f} {g.get('l2_pct', float('nan')):6.1f} "
          f"{g.get('tensor_pct', float('nan')):7.1f} {g.get('sm_pct', float('nan')):6.1f} {g.get('occ_pct', float('nan')):6.1f} {g.get('regs', 0):5.0f} "
          f"{g.get('grid', 0):6.0f}x{g.get('block', 0):<4.0f}")
