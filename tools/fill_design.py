#!/usr/bin/env python
"""Fills the two measured tables of DESIGN.md (between the markers) from bench lines committed under profiles/.
usage: fill_design.py <tag>   (reads profiles/<tag>_bench.json, <tag>_bench_c3.json, <tag>_bench_ref.json)"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]


def line(name):
    for l in open(os.path.join(ROOT, "profiles", name)):
        if l.startswith("{"):
            return json.loads(l)
    raise SystemExit("no JSON line in " + name)


c2, c3, ref = line(f"{tag}_bench.json"), line(f"{tag}_bench_c3.json"), line(f"{tag}_bench_ref.json")


def roof(d, nq):
    r = d["roofline"]
    t = r.get("traffic")
    return (f"| {d['value'] / 1e9:.1f} | {r['kernel_ms']:.3f} | {r['achieved']:.0f} | {r['frac']:.2f} | "
            f"{(t / 1e9):.2f} GB/launch = {t / nq:.0f} B/query | {r['dram_gbs_from_traffic'] / 1e3:.2f} TB/s |" if t else
            f"| {d['value'] / 1e9:.1f} | {r['kernel_ms']:.3f} | {r['achieved']:.0f} | {r['frac']:.2f} | — | — |")


t1 = ["| workload | whole step, G q/s | kernel Q ms/launch | achieved GB/s (algorithmic bytes) | frac of measured HBM copy "
      f"({c2['roofline']['peak']:.0f} GB/s) | DRAM traffic of Q (ncu) | DRAM rate of Q |",
      "|---|---:|---:|---:|---:|---:|---:|",
      "| c2, 1×B200 " + roof(c2, 50e6), "| c3, 1×B200 " + roof(c3, 250e6)]


def steproof(d):
    r = d.get("roofline_step")
    return f"{r['bytes_per_query']:.0f} B/query, {r['achieved']:.0f} GB/s = {r['frac']:.2f} of the copy bandwidth" if r else "—"


def stages(d):
    s = d["roofline"]["stage_ms"]
    return f"{s['hist_scan']:.2f} / {s['scatter']:.2f} / {s['query_kernel']:.2f} / {s['unpermute']:.2f}"


def par(d):
    p = d["parity"]
    k = [x for x in p if x.startswith("mismatches")][0]
    return f"{p['checked'] - p[k]:,} / {p['checked']:,} identical ({k.replace('mismatches_vs_', 'vs the ')})"


t2 = ["| | c2 | c3 |", "|---|---:|---:|",
      f"| GPU device-resident, whole step | {c2['value'] / 1e9:.1f} G q/s ({c2['ms_per_step']:.3f} ms / 50 M) | "
      f"{c3['value'] / 1e9:.1f} G q/s ({c3['ms_per_step']:.2f} ms / 250 M) |",
      f"| stages A+S / B / Q / U, ms | {stages(c2)} | {stages(c3)} |",
      f"| whole-step HBM bound (`roofline_step`) | {steproof(c2)} | {steproof(c3)} |",
      f"| partition bits | {c2['roofline']['partition_bits']} | {c3['roofline']['partition_bits']} |",
      f"| GPU end-to-end (host buffers, PCIe inside) | {c2['e2e']['value'] / 1e9:.2f} G q/s | {c3['e2e']['value'] / 1e9:.2f} G q/s |",
      f"| CPU, {c2['cpu_baseline']['cores']} threads | {c2['cpu_baseline']['value'] / 1e6:.1f} M q/s (unmodified reference; "
      f"`--impl reference`: {ref['value'] / 1e6:.1f}) | {c3['cpu_baseline']['value'] / 1e6:.1f} M q/s (oracle port) |",
      f"| parity at full size | {par(c2)}, `.sap` bytes identical | {par(c3)} |",
      f"| device self-check (`sapling_example.cpp:144-154`) | {c2['self_check']['matching']:,} / {c2['self_check']['of']:,} | "
      f"{c3['self_check']['matching']:,} / {c3['self_check']['of']:,} |",
      f"| SM clock under load | {c2['clocks']['sm_mhz']:.0f} MHz, no throttle reasons | {c3['clocks']['sm_mhz']:.0f} MHz |"]

p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
s = re.sub(r"<!-- ROOFLINE -->.*?<!-- /ROOFLINE -->", "<!-- ROOFLINE -->\n" + "\n".join(t1) + "\n<!-- /ROOFLINE -->", s, flags=re.S)
s = re.sub(r"<!-- BENCH -->.*?<!-- /BENCH -->", "<!-- BENCH -->\n" + "\n".join(t2) + "\n<!-- /BENCH -->", s, flags=re.S)
s = re.sub(r"profiles/r2[g-z]_", f"profiles/{tag}_", s)  # r2f stays: the "before" captures
open(p, "w").write(s)
print("\n".join(t1 + [""] + t2))
