#!/usr/bin/env python3
"""Markdown summary of an `ncu --set full` report: one row per captured launch with the metrics the
roofline discussion uses.  usage: ncu_summary.py <report.ncu-rep> [title]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def g(d, name, default="-"):
    i = col.get(name)
    return d[i] if i is not None and d[i] != "" else default


def f(d, name):
    try:
        return float(g(d, name).replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(d, name):
    i = col.get(name)
    if i is None:
        return float("nan")
    u = units[i].lower()
    v = f(d, name)
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def to_ms(d, name):
    i = col[name]
    u = units[i].lower()
    return f(d, name) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)


print(f"# {title}\n")
print("`ncu --set full --clock-control none`; per launch. tensor% = sm__pipe_tensor_cycles_active (of active cycles); "
      "L2->SM = l1tex__m_xbar2l1tex_read_bytes / duration; DRAM = dram__bytes_read+write.\n")
print("| # | kernel | grid | ms | SM GHz | tensor % | SM active % | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2->SM GB | L2->SM TB/s | regs | smem KB |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, d in enumerate(data):
    name = g(d, "Kernel Name").replace("ecseg::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0]
    ms = to_ms(d, "gpu__time_duration.sum")
    rd, wr = to_bytes(d, "dram__bytes_read.sum"), to_bytes(d, "dram__bytes_write.sum")
    x2l = to_bytes(d, "l1tex__m_xbar2l1tex_read_bytes.sum")
    act = 100.0 * f(d, "sm__cycles_active.avg") / f(d, "sm__cycles_elapsed.avg")
    smem = to_bytes(d, "launch__shared_mem_per_block_dynamic") / 1024
    print(f"| {n} | {name} | {g(d, 'Grid Size')} | {ms:.3f} | {f(d, 'sm__cycles_elapsed.avg.per_second'):.2f} | "
          f"{f(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {act:.0f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | "
          f"{(rd + wr) / ms / 1e6:.0f} | {x2l / 1e9:.2f} | {x2l / ms / 1e9:.2f} | {g(d, 'launch__registers_per_thread')} | {smem:.0f} |")
