#!/usr/bin/env python
"""gpurun_out/r2_gather_probe_{plain,l2_32,l2_64}.log + r2_gather_probe_ncu.csv -> profiles/r2_gather_probe.md
(times from the plain runs, 3 launches each; DRAM bytes per launch from the ncu pass of the same binary)."""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, "gpurun_out")
N = 1 << 28


def rows(path):
    out = []
    for ln in open(path):
        if ln.startswith("ROW "):
            parts = ln.split()
            out.append((parts[1], dict(p.split("=", 1) for p in parts[2:])))
    return out


def ncu_launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    per = collections.OrderedDict()
    for x in csv.DictReader(lines):
        per.setdefault((int(x["ID"]), x["Kernel Name"]), {})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    return [(k[1], v) for k, v in per.items() if not k[1].startswith("fill")]


plain = rows(os.path.join(G, "r2_gather_probe_plain.log"))
launches = ncu_launches(os.path.join(G, "r2_gather_probe_ncu.csv"))
# the ncu run used reps=1: every ROW is (1 warm-up + 1 timed) launches of the same kernel -> take the second of each pair
per_row = [launches[2 * i + 1][1] for i in range(len(launches) // 2)]
assert len(per_row) == len(plain), (len(per_row), len(plain))
md = ["# Round 2 — what a uniform-random gather costs on a B200 (`scripts/probe/gather_probe.cu`)", "",
      "2^28 indices (8 B each, read coalesced), 4-byte elements gathered from a source of the given span, 4 B written per element, V = 4 gathers in flight",
      "per thread.  `ms` and `G gathers/s`: plain runs, CUDA events, mean of 3 launches.  `DRAM B/elt`: `dram__bytes_read.sum` and",
      "`dram__bytes_write.sum` of one launch from `ncu --metrics … --clock-control none` on the same binary, divided by 2^28.", "",
      "| source span | load flavour | V | ms | G gathers/s | algorithmic GB/s (16 B/elt) | DRAM read B/elt | DRAM write B/elt |", "|---|---|---|---|---|---|---|---|"]
extra = []
for (kind, f), m in zip(plain, per_row):
    rd, wr = m["dram__bytes_read.sum"] / N, m["dram__bytes_write.sum"] / N
    if kind == "gather":
        span = int(f["span"].split("^")[1])
        md.append(f"| 2^{span} elements ({4 * 2**span / 2**20:.0f} MiB) | `{f['flavour']}` | {f['V']} | {f['ms']} | {f['ggather_s']} | {f['alg_gbs']} | {rd:.1f} | {wr:.1f} |")
    else:
        extra.append((kind, f, rd, wr))
md += ["", "## The halves of a partitioned (two-pass) gather", "",
       "Indices pre-bucketed by source slab (so that the slab being read is L2-resident), 2^28 elements, 4 GiB source:", "",
       "| step | slab | ms | G elements/s | DRAM read B/elt | DRAM write B/elt |", "|---|---|---|---|---|---|"]
for kind, f, rd, wr in extra:
    slab = f.get("slab", "—")
    rate = f.get("ggather_s", f.get("gscatter_s"))
    md.append(f"| `{kind}` | {slab} | {f['ms']} | {rate} | {rd:.1f} | {wr:.1f} |")
for tag in ("l2_32", "l2_64"):
    p = os.path.join(G, f"r2_gather_probe_{tag}.log")
    if os.path.exists(p):
        r = [x for x in rows(p) if x[0] == "gather" and x[1]["flavour"] == "ldg" and x[1]["V"] == "4" and x[1]["span"] == "2^30"]
        if r:
            md.append("")
            md.append(f"`cudaLimitMaxL2FetchGranularity = {tag.split('_')[1]}`: 4 GiB span, `ldg`, V=4: {r[0][1]['ms']} ms ({r[0][1]['ggather_s']} G gathers/s) — no change.")
md += ["", "## Reading", "",
       "* A random 4-byte read of a 4 GiB source moves **one whole 128-byte line from DRAM** (134 B/elt read = 8 B index + 126 B): 36 GB of DRAM",
       "  traffic for 4.3 GB of algorithmic bytes, at 44 G lines/s = 5.7 TB/s of DRAM reads.  Load flavours (`.nc`, `.cv`, `.cs`, `.lu`,",
       "  `L1::no_allocate`), the number of gathers in flight per thread and `cudaLimitMaxL2FetchGranularity` change nothing.",
       "* `ld.global.nc.L2::64B` HALVES the DRAM bytes (72 B/elt) but the time only drops 6 % (5.66 ms): the limit is the **rate of random DRAM",
       "  accesses (~45-47 G/s), not bytes** — row activations, not bandwidth.  The product's gather uses this flavour when the gathered source",
       "  spans more than 1 GiB (`PF_GATHER_BIG`, csrc/exec.cuh `ld32_big`); below that the plain load wins because part of the source stays in L2",
       "  (256 MiB: 4.25 vs 4.47 ms; 64 MiB: 1.23 vs 1.27 ms).",
       "* Once the source fits L2 (<= 64 MiB) the same kernel gathers 5x faster (219 G/s, 18 B/elt of DRAM traffic): that is what a partitioned gather",
       "  would buy on the READ side (`bucketed_gather_ordered_out`: 1.56 ms).",
       "* But the results then have to go back to their positions, and **random 4-byte scatters are twice as slow as random gathers** (23.9 G/s:",
       "  each one is a 32-byte sector read-modify-write, 35 B read + 32 B written per element): gather-from-L2 + scatter = 13.1 ms against 6.0 ms for",
       "  the direct gather.  Making the write side local as well needs a second partition pass; with the measured rates the whole pipeline",
       "  (histogram 0.33 + partition 0.65 + gather 1.6 + histogram 0.33 + partition 0.65 + in-L2 scatter ~0.8 ms) comes to ~4.4 ms at best against",
       "  5.66 ms for the direct gather with 64-byte fetches, for ~8 GiB of scratch memory and six kernels.  Not built: the one-pass gather with",
       "  `L2::64B` already brings total DRAM traffic from 36.9 GB to 20.4 GB (the figure the partitioned variant was meant to reach).",
       "* So config 3 is bound by the DRAM random-access rate: 16 algorithmic bytes per element x 47 G/s = 0.76 TB/s = 0.095 of 8 TB/s.  The",
       "  north star's 75 % bar is not reachable for uniform-random 4-byte indices on this part by any one- or two-pass scheme measured here."]
open(os.path.join(ROOT, "profiles", "r2_gather_probe.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[:14]))
