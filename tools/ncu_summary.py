"""Summarise ncu output into profiles/: (a) the per-launch list (gpu__time_duration) aggregated by kernel,
(b) key metrics of every kernel in one or more `--set full` reports.
    python tools/ncu_summary.py --launches gpurun_out/launches_r01.csv --reports gpurun_out/prof_*.ncu-rep --out profiles/r01_ncu_summary.md
"""
import argparse
import collections
import csv
import io
import re
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
    ("sm__cycles_elapsed.max.per_second", "SM clock"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("rvl::", "")


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
        key = short(row["Kernel Name"]) + " grid=" + row["Grid Size"].replace(" ", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    by_kernel = collections.OrderedDict()
    for k, (n, t) in agg.items():
        b = by_kernel.setdefault(k.split(" grid=")[0], [0, 0.0])
        b[0] += n
        b[1] += t
    out.write(f"## Launch list of one sweep ({sum(n for n, _ in agg.values())} launches, {tot / 1e3:.1f} ms of kernel time, "
              "cold-cache serialised ncu timing: compare shares)\n\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(by_kernel.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{k}` | {n} | {t / 1e3:.2f} | {100 * t / tot:.1f}% | {t / n:.1f} |\n")
    out.write("\n### Same list split by grid size (top 25)\n\n| kernel, grid | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        out.write(f"| `{k}` | {n} | {t / 1e3:.2f} | {100 * t / tot:.1f}% | {t / n:.1f} |\n")
    out.write("\n")


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out.write(f"## `{path.split('/')[-1]}` (ncu --set full --clock-control none)\n\n")
    out.write("| # | kernel | grid | " + " | ".join(lbl for _, lbl in KEYS) + " |\n|---|---|---|" + "---:|" * len(KEYS) + "\n")
    for i, r in enumerate(rows[2:]):
        cells = []
        for k, _ in KEYS:
            if k in col:
                v = r[col[k]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[col[k]]}".strip())
            else:
                cells.append("-")
        out.write(f"| {i} | `{short(r[col['Kernel Name']])}` | {r[col['Grid Size']]} | " + " | ".join(cells) + " |\n")
    out.write("\n")


def traffic(path, out_json, algorithmic_bytes):
    """Average measured DRAM bytes (read + write) per launch over the kernels of one `--set full` report -> the
    `roofline.traffic` value bench.py reports for the dominant kernel."""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = []
    for r in rows[2:]:
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[k]].replace(",", "")) * scale[units[col[k]]]
        per.append(b)
    json.dump({"prefill_gemm_dram_bytes_per_launch": sum(per) / len(per), "per_launch": per,
               "prefill_gemm_algorithmic_bytes_per_launch": algorithmic_bytes,
               "source": f"ncu --set full --clock-control none, {path.split('/')[-1]}: the 4 prefill GEMMs of one layer (qkv, o, gate|up+SwiGLU, down), "
                         "dram__bytes_read.sum + dram__bytes_write.sum averaged per launch"}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--reports", nargs="*", default=[])
    ap.add_argument("--out", required=True)
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--traffic-report", help="--set full report of the dominant kernel's launches")
    ap.add_argument("--traffic-out")
    ap.add_argument("--algorithmic-bytes", type=float, default=0.0)
    a = ap.parse_args()
    if a.traffic_report and a.traffic_out:
        traffic(a.traffic_report, a.traffic_out, a.algorithmic_bytes)
    with open(a.out, "w") as f:
        f.write(f"# {a.title}\n\n")
        if a.launches:
            launches(a.launches, f)
        for r in a.reports:
            report(r, f)
    print(open(a.out).read())
