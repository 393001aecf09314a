"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_bench_launches.md "<command>"
    python tools/summarize_profiles.py full gpurun_out/mlp_r1.ncu-rep profiles/r1_mlp_ncu_full.md "<command>"
"""
import csv
import collections
import io
import subprocess
import sys


def launches(src, dst, cmd):
    rows = [l for l in open(src) if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = r["Kernel Name"].split("(")[0]
        a = per.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\ncommand: `{cmd}`\n\n")
        f.write(f"{n} launches captured, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t:.1f} | {100 * t / total:.2f} % | {t / c:.2f} |\n")
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg ", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg ", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum ", "sm__inst_executed_pipe_uniform", "launch__occupancy_limit",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct"]


def full(src, dst, cmd):
    out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary\n\ncommand: `{cmd}`\n\nsource report: `{src}` (scratch, not tracked)\n\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name}\n\n| metric | unit | value |\n|---|---|---:|\n")
            for h, u, v in zip(hdr, units, vals):
                if any((h + " ").startswith(w) or h == w.strip() for w in WANT):
                    f.write(f"| {h} | {u} | {v} |\n")
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
