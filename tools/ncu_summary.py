#!/usr/bin/env python
"""Turns the ncu artefacts a GPU session left in gpurun_out/ into the small text summaries committed under
profiles/ (runs here, on the CPU box: `ncu -i` needs no GPU).

  python tools/ncu_summary.py launches gpurun_out/<tag>_launches.csv profiles/<tag>_launches.md
  python tools/ncu_summary.py full     gpurun_out/<tag>_query.ncu-rep profiles/<tag>_query_ncu.md [queries_per_launch]
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("sb::<unnamed>::", "")
    name = re.sub(r"<.*", "", name)
    return name


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    kn, val, grid, block = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    seq = []
    for r in rows:
        ns = float(r[val].replace(",", ""))
        k = short(r[kn])
        a = agg.setdefault(k, [0, 0.0, r[grid], r[block]])
        a[0] += 1
        a[1] += ns
        seq.append((r[0], k, ns, r[grid], r[block]))
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over one `bench.py` run (index build, "
                "query sampling, warm-up + timed query steps, verification, e2e).  Times are cold-cache and serialised: "
                "compare SHARES.\n\n")
        f.write(f"{len(rows)} launches, {total / 1e6:.3f} ms of device time in total.\n\n")
        f.write("| kernel | launches | total ms | share | mean us | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.1f}% | {a[1] / a[0] / 1e3:.1f} | {a[2]} | {a[3]} |\n")
        q = [s for s in seq if "kmer_query" in s[1]]
        if q:
            f.write("\nQuery-kernel launches in order (ns): " + ", ".join(f"{int(s[2])}" for s in q) + "\n")
            steady = [s for s in seq if any(t in s[1] for t in ("kmer_query", "verify", "sample"))]
            st = sum(s[2] for s in steady)
            f.write(f"\nShare of the query kernel within the query phase (query + sample + verify launches): "
                    f"{100 * sum(s[2] for s in q) / st:.1f}%\n")


KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__warps_active.avg.per_cycle_active", None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__sectors_read.sum", None),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__bytes.sum.per_second", "DRAM bytes/s"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (LTS) throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs (both dies' lookups)"),
    ("lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", None),
    ("lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", None),
    ("lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_hit.sum", "  evict_last (genome+model) lookup hits"),
    ("lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_miss.sum", "  evict_last lookup misses"),
    ("lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_hit.sum", "  evict_first (SA, k-mers) lookup hits"),
    ("lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_miss.sum", "  evict_first lookup misses"),
    ("lts__t_requests_srcunit_ltcfabric.sum", "L2 die-to-die fabric requests"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global-load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum", "sectors L2 -> L1"),
    ("l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ldgsts.sum", "sectors L2 -> smem (cp.async)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "bytes L2 -> SM"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per issued instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "  stalled on long scoreboard (memory)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def full(src, dst, nq=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n`ncu --set full --clock-control none --import-source on -k regex:kmer_query "
                "-s 3 -c 1 python bench.py --steps 3 --warmup 3 --cpu-baseline none`; read with `ncu -i ... --page raw --csv`.\n")
        for vals in rows[2:]:
            d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
            f.write(f"\n## {short(d['Kernel Name'][1])}  grid {d.get('Grid Size', ('', ''))[1]} block {d.get('Block Size', ('', ''))[1]}\n\n")
            f.write("| metric | value | unit | per query |\n|---|---:|---|---:|\n")
            for key, label in KEYS:
                if key not in d:
                    continue
                u, v = d[key]
                per = ""
                try:
                    x = float(v.replace(",", ""))
                    if nq and (key.endswith(".sum") and ("sector" in key or "bytes" in key or "request" in key or "inst" in key)):
                        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(u, 1)
                        per = f"{x * scale / nq:.2f}"
                except ValueError:
                    pass
                f.write(f"| {label or key} (`{key}`) | {v} | {u} | {per} |\n")
        # stall profile of the first kernel from the source page
        srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        lines = list(csv.reader(io.StringIO(srcp)))
        h = None
        for i, l in enumerate(lines):
            if l and l[0] == "Address":
                h = i
                break
        if h is not None:
            hd = lines[h]
            si, ii = hd.index("Source"), hd.index("# Samples")
            body = []
            for l in lines[h + 1:]:  # a report with several kernels repeats the header: keep the first kernel only
                if l and l[0] == "Address":
                    break
                if len(l) == len(hd):
                    body.append(l)
            tot = sum(int(l[ii] or 0) for l in body)
            top = sorted(body, key=lambda l: -int(l[ii] or 0))[:14]
            f.write(f"\n## Hottest SASS instructions by warp-stall samples ({tot} samples)\n\n| samples | share | instruction |\n|---:|---:|---|\n")
            for l in top:
                f.write(f"| {l[ii]} | {100 * int(l[ii]) / max(tot, 1):.1f}% | `{l[si].strip()}` |\n")
            stall_cols = [c for c in hd if c.startswith("stall_") and "Not Issued" not in c]
            sums = {c: sum(int(l[hd.index(c)] or 0) for l in body) for c in stall_cols}
            f.write("\nStall reasons (all samples): " + ", ".join(f"{c[6:]} {100 * v / max(tot, 1):.1f}%" for c, v in
                                                                sorted(sums.items(), key=lambda kv: -kv[1]) if v) + "\n")


if __name__ == "__main__":
    what, src, dst = sys.argv[1:4]
    if what == "launches":
        launches(src, dst)
    else:
        full(src, dst, float(sys.argv[4]) if len(sys.argv) > 4 else None)
