#!/usr/bin/env python
"""Turn ncu output into the tracked summaries under profiles/.

  python profiles/summarize_ncu.py full  gpurun_out/prof_X.ncu-rep  profiles/rN_ncu_full.md  "<command that was profiled>"
      -> one table per captured launch with the metrics the roofline argument uses (needs `ncu` on PATH to
         read the report: `ncu -i ... --page raw --csv`); with --traffic OUT.json KERNEL_REGEX ALG_BYTES also
         writes dram bytes per launch of the LONGEST matching launch (bench.py's roofline.traffic).
  python profiles/summarize_ncu.py list  gpurun_out/launches.csv    profiles/rN_launches_summary.md "<command>"
      -> per-kernel launch counts / total / average of a `--metrics gpu__time_duration.sum --csv` launch list.
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"gf::\(anonymous namespace\)::|gf::<unnamed>::|\(anonymous namespace\)::|<unnamed>::", "", name)
    return name if len(name) < 110 else name[:107] + "..."


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1.0)


def read_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    return header, units, data


def cmd_full(argv):
    rep, out_md, command = argv[0], argv[1], argv[2]
    traffic = None
    if "--traffic" in argv:
        i = argv.index("--traffic")
        traffic = (argv[i + 1], argv[i + 2], int(argv[i + 3]))
    header, units, data = read_raw(rep)
    col = {h: i for i, h in enumerate(header)}
    lines = ["# ncu --set full (1x B200)", "", "Command: `%s` (after the same command exited 0 without ncu); read with "
             "`ncu -i %s --page raw --csv` by profiles/summarize_ncu.py." % (command, rep.split("/")[-1]), ""]
    best = None
    for row in data:
        name = row[col["Kernel Name"]]
        lines += ["## " + short(name), "", "| metric | value |", "|---|---|"]
        for m in KEEP:
            if m in col and row[col[m]] != "":
                lines.append("| %s | %s %s |" % (m, row[col[m]], units[col[m]]))
        lines.append("")
        if traffic and re.search(traffic[1], name):
            dur = float(row[col["gpu__time_duration.sum"]].replace(",", ""))
            rd = to_bytes(row[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(row[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            if best is None or dur > best[0]:
                best = (dur, name, rd, wr)
    with open(out_md, "w") as f:
        f.write("\n".join(lines))
    if traffic and best:
        with open(traffic[0], "w") as f:
            json.dump({"kernel": short(best[1]), "dram_bytes_per_launch": best[2] + best[3], "dram_bytes_read": best[2],
                       "dram_bytes_write": best[3], "algorithmic_bytes": traffic[2],
                       "source": "%s (ncu --set full, 1 launch)" % out_md}, f, indent=1)


def cmd_list(argv):
    src, out_md, command = argv[0], argv[1], argv[2]
    text = open(src).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r.get("Metric Unit") == "ns" else v
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    lines = ["# ncu launch list (1x B200)", "", "Command: `%s`, right after the same command exited 0 without ncu. "
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes. Raw list: `%s`."
             % (command, src.split("/")[-1]), "", "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f | %.1f %% |" % (k, n, t, t / n, 100 * t / total))
    lines.append("")
    with open(out_md, "w") as f:
        f.write("\n".join(lines))


if __name__ == "__main__":
    {"full": cmd_full, "list": cmd_list}[sys.argv[1]](sys.argv[2:])
