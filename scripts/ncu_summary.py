#!/usr/bin/env python
"""Turn ncu outputs into the small tracked summaries under profiles/.
  python scripts/ncu_summary.py full  <file.ncu-rep> <out.md>      # --set full capture: one row per launch
  python scripts/ncu_summary.py launches <launches.csv> <out.md>   # gpu__time_duration list: share per kernel
"""
import csv, io, subprocess, sys, collections, re

def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|mdim::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]

def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp instrs"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %")]
    with open(out, "w") as f:
        f.write(f"ncu --set full --clock-control none, source: {rep.split('/')[-1]} (cold-cache, serialised replays; see bench.py for timed numbers)\n\n")
        f.write("| # | kernel | " + " | ".join(f"{t} [{units[idx[m]]}]" if units[idx[m]] else t for m, t in cols if m in idx) + " |\n")
        f.write("|---|---|" + "---|" * sum(1 for m, _ in cols if m in idx) + "\n")
        for k, r in enumerate(rows[2:]):
            vals = []
            for m, _ in cols:
                if m not in idx: continue
                v = r[idx[m]]
                try: v = f"{float(v):.4g}"
                except ValueError: pass
                vals.append(v)
            f.write(f"| {k} | `{short(r[idx['Kernel Name']])}` | " + " | ".join(vals) + " |\n")

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        name, ns = short(r[4]), float(r[-1])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns; total += ns
    with open(out, "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, source: {path.split('/')[-1]}; {len(rows)} launches, {total/1e6:.2f} ms of kernel time "
                f"(cold-cache, serialised: compare SHARES, not absolutes)\n\n| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns/1e6:.3f} | {ns/n/1e3:.2f} | {100*ns/total:.1f}% |\n")

if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
