"""Turns the ncu outputs of scripts/r01_final_measure.sh into the tracked summaries under profiles/.
usage: python scripts/ncu_summaries.py launches <launches.csv> <out.md> "<command>"
       python scripts/ncu_summaries.py full <raw.csv>[,<raw2.csv>...] <out.md> "<command>"     (ncu -i x.ncu-rep --page raw --csv)"""
import csv, re, sys
from collections import OrderedDict

def short(name):
    name = re.sub(r"sonar::\(anonymous namespace\)::|sonar::<unnamed>::|<unnamed>::|unnamed>::|\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()

def launches(path, out, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e6  # ns -> ms
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list of `{cmd}` (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Cold-cache, serialised per-launch times (ncu replays every kernel alone, so the overlap of the alignment "
                "branch with the YIN kernel is not visible here): compare SHARES with the live CUDA-event shares in the bench "
                "JSON (`kernels`).\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.3f} | {ms / tot:.3f} |\n")

M = [("time ms", "gpu__time_duration.sum", "time"), ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1),
     ("regs", "launch__registers_per_thread", 1), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
     ("issue/cycle/SMSP", "smsp__issue_active.avg.per_cycle_active", 1),
     ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
     ("fma pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
     ("l1tex %", "l1tex__throughput.avg.pct_of_peak_sustained_active", 1),
     ("smem wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
     ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
     ("dram read MB", "dram__bytes_read.sum", None), ("dram write MB", "dram__bytes_write.sum", None),
     ("warp inst", "smsp__inst_executed.sum", 1)]

def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

def full(paths, out, cmd):
    cols = []
    for path in paths.split(","):
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            g = lambda name: (r[hdr.index(name)].replace(",", ""), units[hdr.index(name)])
            d = {"name": short(r[hdr.index("Kernel Name")])}
            for label, metric, scale in M:
                v, u = g(metric)
                v = float(v)
                if scale == "time":
                    d[label] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3,
                                    "msecond": 1.0, "second": 1e3}[u]
                else:
                    d[label] = to_bytes(v, u) / 1e6 if scale is None else v * scale
            st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[i].replace(",", "")) for i, h in enumerate(hdr)
                  if "smsp__average_warps_issue_stalled_" in h and h.endswith("per_issue_active.ratio")}
            d["stalls"] = sorted(st.items(), key=lambda kv: -kv[1])[:4]
            cols.append(d)
    with open(out, "w") as f:
        f.write(f"# ncu --set full, one launch of each kernel of the step\n\nCommand: `{cmd}`\n\n")
        f.write("| metric | " + " | ".join(f"`{c['name']}`" for c in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        for label, _, _ in M:
            f.write(f"| {label} | " + " | ".join(f"{c[label]:.4g}" for c in cols) + " |\n")
        f.write("\nTop issue-stall reasons (warps stalled per issue):\n\n")
        for c in cols:
            f.write(f"* `{c['name']}`: " + ", ".join(f"{k} {v:.2f}" for k, v in c["stalls"]) + "\n")

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
