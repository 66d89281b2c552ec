#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small, tracked summaries under profiles/.

  python benchmarks/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_bench_launches
  python benchmarks/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_top_kernels
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_tma_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("usf::", "").replace("<unnamed>::", "").replace("unnamed>::", "")


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
    out, agg, tot = [], collections.OrderedDict(), 0.0
    for r in rows:
        us = float(r["Metric Value"]) / 1e3
        k = short(r["Kernel Name"])
        out.append((r["ID"], k, r["Grid Size"], r["Block Size"], us))
        fam = k.split("<")[0]
        agg.setdefault(fam, [0, 0.0])
        agg[fam][0] += 1
        agg[fam][1] += us
        tot += us
    with open(dst + ".csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu_time_us"])
        w.writerows(out)
    with open(dst + ".md", "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none (cold cache, serialised): {len(out)} launches, "
                f"{tot:.1f} us total.  Compare SHARES, not absolutes.\n\n| kernel family | launches | us | share |\n|---|---|---|---|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {us:.1f} | {100 * us / tot:.1f}% |\n")
    print(open(dst + ".md").read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    res = []
    for d in data:
        e = {"kernel": short(d[ix["Kernel Name"]])}
        for k in KEYS:
            if k in ix:
                e[k] = f"{d[ix[k]]} {units[ix[k]]}".strip()
        st = sorted(((float(d[ix[h]] or 0), h) for h in stall), reverse=True)[:5]
        e["top_stalls_per_issue"] = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 2)
                                     for v, h in st}
        res.append(e)
    with open(dst + ".json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
