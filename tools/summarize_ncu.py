"""Turn ncu outputs into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md [launches_per_step]
    python tools/summarize_ncu.py full gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_kernels.md
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(unsigned char\)|\(int\)|\(bool\)", "", name)
    name = name.replace("b200::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"^unnamed>::", "", name)
    m = re.match(r"(\w+)(<.*>)?\(", name)
    if m:
        return m.group(1) + (m.group(2) or "")
    m = re.match(r"(\w+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:80]


def launches(path, out, per_step=None):
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
    if per_step:
        rows = rows[-int(per_step):]
    agg = collections.OrderedDict()
    for d in rows:
        t = float(d["Metric Value"].replace(",", ""))
        t = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}.get(d["Metric Unit"], t)
        k = short(d["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path}), last {len(rows)} launches; cold-cache serialised times: compare SHARES\n\n")
        f.write(f"total {tot:.1f} us over {len(rows)} launches\n\n| share | us | launches | kernel | grid | block |\n|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * a[1] / tot:.1f}% | {a[1]:.1f} | {a[0]} | `{k}` | {a[2]} | {a[3]} |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({path}); one block per captured launch\n\n")
        for d in data:
            f.write(f"## `{short(d[hdr.index('Kernel Name')])}` grid {d[hdr.index('Grid Size')]} block {d[hdr.index('Block Size')]}\n\n")
            for m, i in cols:
                f.write(f"- {m}: {d[i]} {units[i]}\n")
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        full(sys.argv[2], sys.argv[3])
