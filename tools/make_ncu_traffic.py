"""profiles/ncu_traffic.json: measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the tensor-core
convolution kernels, keyed the way bench.py names its op launches, from `ncu --set full` raw CSV pages.

    python tools/make_ncu_traffic.py profiles/ncu_traffic.json <raw.csv>:<per_gpu_batch> [<raw.csv>:<per_gpu_batch> ...]
Entries are scaled linearly to per-GPU batch 1024 (all tensors scale with the batch)."""
import csv
import json
import re
import sys

V_STUDENT, V_TEACHER = 6, 2


def col(hdr, name):
    if name in hdr:
        return hdr.index(name)
    for i, h in enumerate(hdr):
        if h.endswith(name):
            return i
    raise KeyError(name)


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    out, specs = sys.argv[1], sys.argv[2:]
    launches = {}
    for spec in specs:
        path, b = spec.rsplit(":", 1)
        b = int(b)
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ik, ig = hdr.index("Kernel Name"), hdr.index("Grid Size")
        it, ir, iw = col(hdr, "gpu__time_duration.sum"), col(hdr, "dram__bytes_read.sum"), col(hdr, "dram__bytes_write.sum")
        itc = col(hdr, "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed")
        il1 = col(hdr, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")
        for d in data:
            name = re.sub(r"\(int\)|\(bool\)", "", d[ik])
            grid = [int(x) for x in re.findall(r"\d+", d[ig])]
            traffic = (to_bytes(d[ir], units[ir]) + to_bytes(d[iw], units[iw])) * 1024 / b
            def qw(w, pad):
                return (w + 2 * pad + 3) // 4
            m = re.search(r"conv_tc_kernel<.*TcCfg<([\d, ]+)>", name)
            m2 = re.search(r"conv_tc_wgrad_ph_kernel<.*WgPCfg<([\d, ]+)>", name)
            m3 = re.search(r"conv_tc_wgrad_l0_fused_kernel<.*L0FCfg<([\d, ]+)", name)
            m4 = re.search(r"conv_tc_wgrad_kernel<.*TcWgCfg<([\d, ]+)>", name)
            if m:
                a = [int(x) for x in m.group(1).split(",")]
                cin, cout, hin, win, ks, pad = a[0], a[1], a[3], a[4], a[5], a[6]
                views = grid[1]
                n = 1024 * (views if views > 1 else V_STUDENT)
                npv = 1024 if views > 1 else n
                shape = f"{n}x{hin}x{qw(win, pad)}x8" if cin == 1 else f"{n}x{cin // 8}x{hin}x{win}x8"
                pooled = re.search(r">\s*,\s*(1|true)\s*>", name) is not None        # conv_tc_kernel<TcCfg<...>, POOL>
                key = f"{'conv_tc_pool' if pooled else 'conv_tc'}:{shape}:{npv}x{cout}x{ks}x{pad}"
            elif m2:
                a = [int(x) for x in m2.group(1).split(",")]
                cin, hin, win, pad = a[0], a[2], a[3], a[5]
                key = f"conv_tc_wgrad:{1024 * V_STUDENT}x{cin // 8}x{hin}x{win}x8:{pad}"
            elif m3:
                a = [int(x) for x in m3.group(1).split(",")]
                hin, win, pad = a[1], a[2], a[4]
                key = f"conv_tc_wgrad_l0_fused:{1024 * V_STUDENT}x{hin}x{qw(win, pad)}x8:1024x{pad}"
            elif m4:
                a = [int(x) for x in m4.group(1).split(",")]
                cin, hin, win, pad = a[0], a[2], a[3], a[5]
                n = 1024 * V_STUDENT
                shape = f"{n}x{hin}x{win + pad}x8" if cin == 1 else f"{n}x{cin // 8}x{hin}x{win}x8"
                key = f"conv_tc_wgrad:{shape}:{pad}"
            else:
                continue
            if key in launches:           # the same kernel more than once in the capture: keep the first (student) launch
                continue
            launches[key] = {"tensor_pipe_busy_pct": float(d[itc] or 0), "l1tex_pct": float(d[il1] or 0), "dram_bytes": traffic, "us_under_ncu": float(d[it].replace(",", "")) * (1024 / b if False else 1), "source": path.split("/")[-1],
                             "captured_at_batch": b}
    json.dump({"per_gpu_batch": 1024, "metric": "dram__bytes_read.sum + dram__bytes_write.sum (ncu --set full --clock-control none)",
               "launches": launches}, open(out, "w"), indent=1)
    print(len(launches), "launches")


if __name__ == "__main__":
    main()
