#!/usr/bin/env python
"""Digest of an .ncu-rep file read on the build box: per kernel the headline metrics, the stall-reason
histogram and the hottest source instructions.  python benchmarks/ncu_digest.py file.ncu-rep [--top N] [--kernel regex]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("=" * 100)
        print(name[:140])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>14s} {units[hdr.index(k)]}")
        st = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1
        print("  stalls:", ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:7]))
    # source page (all kernels concatenated): hottest instructions
    src = page(rep, "source")
    cur, block = None, []

    def flush():
        if not block:
            return
        h = block[0]
        si, ii = h.index("# Samples"), h.index("Instructions Executed")
        body = [b for b in block[1:] if len(b) > si and b[si].isdigit()]
        tot = sum(int(b[si]) for b in body) or 1
        toti = sum(int(b[ii]) for b in body)
        print("-" * 100)
        print(cur[:120], f"| samples {tot}, warp instructions {toti}")
        for b in sorted(body, key=lambda b: -int(b[si]))[:top]:
            print(f"   {100 * int(b[si]) / tot:5.1f}%  x{b[ii]:>9s}  {b[1].strip()[:100]}")

    for r in src:
        if r and r[0] == "Kernel Name":
            flush()
            cur, block = r[1], []
        elif r and r[0] == "Address":
            block = [r]
        elif block:
            block.append(r)
    flush()


if __name__ == "__main__":
    main()
