"""Turn ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md "title"
    python tools/summarize_ncu.py full gpurun_out/prof_lik.ncu-rep profiles/r01_likelihood_full.md "title"
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    while lines and "Kernel Name" not in lines[0]:
        lines.pop(0)
    agg = collections.OrderedDict()
    order = []
    all_rows = list(csv.DictReader(lines))
    ours = [r for r in all_rows if "scvae" in r["Kernel Name"]]
    for row in (ours or all_rows):   # (a -k filtered capture lists base names without namespace)
        name = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        short = name.split("(")[0].replace("void ", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        order.append((short, v, row["Grid Size"], row["Block Size"]))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as out:
        out.write("# {}\n\n".format(title))
        out.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch "
                  "times are cold-cache and serialised: compare SHARES, not absolutes). Only this "
                  "repository's kernels (`scvae::*`) are listed; total {:.1f} us over {} launches.\n\n"
                  .format(tot, len(order)))
        out.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("| `{}` | {} | {:.1f} | {:.1f} | {:.1f}% |\n".format(k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
        out.write("\n## Launch sequence of the last training step\n\n| # | kernel | us | grid | block |\n|---|---|---:|---|---|\n")
        # last TRAINING step = the last csr_densify that is followed by the optimiser, up to the
        # next csr_densify (evaluation passes have no optimiser launch)
        starts = [i for i, o in enumerate(order) if "densify" in o[0]] + [len(order)]
        idx, end = 0, len(order)
        for a, b in zip(starts[:-1], starts[1:]):
            if any("adam" in o[0] or "dp_reduce" in o[0] for o in order[a:b]):
                idx, end = a, b
        for i, o in enumerate(order[idx:end]):
            out.write("| {} | `{}` | {:.1f} | {} | {} |\n".format(i, o[0], o[1], o[2], o[3]))


def full(src, dst, title):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write("# {}\n\nSource: `ncu --set full --clock-control none --import-source on` ({}).\n\n".format(title, src))
        for r in rows[2:]:
            out.write("## {}\n\n| metric | value | unit |\n|---|---:|---|\n".format(r[hdr.index("Kernel Name")].split("(")[0]))
            for m in FULL_METRICS:
                if m in hdr:
                    out.write("| {} | {} | {} |\n".format(m, r[hdr.index(m)], units[hdr.index(m)]))
            out.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
