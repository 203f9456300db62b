"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

  python scripts/ncu_summaries.py launches gpurun_out/launches12.csv profiles/r1_v5_launches_venice "header text"
  python scripts/ncu_summaries.py full gpurun_out/prof_v12.ncu-rep profiles/r1_v5_kernels_full.txt "header text"
"""
import collections, csv, io, shutil, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]


def launches(src, dst, header):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        a = agg.setdefault(r[kn][:100], [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1000.0
    tot = sum(a[1] for a in agg.values())
    with open(dst + "_summary.txt", "w") as f:
        f.write(header + "\n# per-launch times are cold-cache and serialised: compare SHARES\n\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{a[1]:12.1f} us {100 * a[1] / tot:5.1f}% n={a[0]:4d} avg {a[1] / a[0]:9.1f} us  {n}\n")
    shutil.copy(src, dst + ".csv")


def full(src, dst, header):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, units = r[0], r[1]
    with open(dst, "w") as f:
        f.write(header + "\n")
        for row in r[2:]:
            f.write("===== " + row[h.index("Kernel Name")][:60] + "\n")
            for m in METRICS:
                if m in h:
                    i = h.index(m)
                    f.write(f"   {m} = {row[i]} {units[i]}\n")
            st = []
            for i, c in enumerate(h):
                if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio"):
                    st.append((float(row[i] or 0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            st.sort(reverse=True)
            f.write("   stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:7]) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
