#!/usr/bin/env python
"""Partitioned batch path (partition.cu) against the unpartitioned kernel: same index, same queries, identical answers.
usage: part_sweep.py <genome bp> <queries> <layouts: plain,packed3,packed4,inline> <bits list, 0 = off> [blocks/SM list] [hints list]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sapling_b200 as S

n = int(float(sys.argv[1]))
nq = int(float(sys.argv[2]))
layouts = sys.argv[3].split(",")
bits_list = [int(b) for b in sys.argv[4].split(",")]
qv_list = [int(b) for b in (sys.argv[5] if len(sys.argv) > 5 else "4").split(",")]
hints_list = [int(b) for b in (sys.argv[6] if len(sys.argv) > 6 else "3").split(",")]
st = torch.cuda.current_stream().cuda_stream
d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
rows, ref = [], {}
os.environ["SAPLING_B200_PART_MIN"] = "1"
for lay in layouts:
    os.environ.pop("SAPLING_B200_PACKED_SHIFT", None)
    flags = S.QUIET | {"plain": S.NO_PACKED | S.NO_INLINE, "inline": S.INLINE | S.NO_PACKED,
                       "packed3": S.PACKED | S.NO_INLINE, "packed4": S.PACKED | S.NO_INLINE}[lay]
    if lay == "packed4":
        os.environ["SAPLING_B200_PACKED_SHIFT"] = "4"
    for hints in hints_list:
        os.environ["SAPLING_B200_HINTS"] = str(hints)
        t0 = time.time()
        ix = S.Sapling.synthetic(0x5A911C0DE5EED001, n, k=21, maxMem=10, flags=flags)
        torch.cuda.synchronize()
        build_s = time.time() - t0
        for mut in (0, 0x5A911C0DE5EED003):
            ix.sample_queries_device(0x5A911C0DE5EED002, mut, 0, nq, d_k.data_ptr(), st)
            for bits in bits_list:
                if bits:
                    os.environ["SAPLING_B200_PART"] = "1"
                    os.environ["SAPLING_B200_PART_BITS"] = str(bits)
                else:
                    os.environ["SAPLING_B200_PART"] = "0"
                for qv in qv_list:
                    os.environ["SAPLING_B200_QV"] = str(qv)
                    for _ in range(2):
                        ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(4):
                        ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 4
                    ix.profile(True)  # stage breakdown, separately from the timed loop
                    for _ in range(3):
                        ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
                    calls, stage = ix.stage_ms()
                    ix.profile(False)
                    stage = [round(x / max(calls, 1), 3) for x in stage]
                    if mut not in ref:
                        ref[mut] = d_o.clone()
                    rows.append({"genome_bp": n, "queries": nq, "layout": lay, "kernel": ix.query_kernel()[0], "hints": hints,
                                 "part_bits": bits, "mutated_half": bool(mut), "blocks_per_sm": qv, "ms": round(ms, 3),
                                 "Gq_per_s": round(nq / ms / 1e6, 2), "same_results": bool(torch.equal(d_o, ref[mut])),
                                 "device_MB": round(ix.device_bytes() / 1e6), "build_s": round(build_s, 2),
                                 "stage_ms": dict(zip(("hist_scan", "scatter", "query", "unpermute"), stage)),
                                 "scatter": os.environ.get("SAPLING_B200_PART_SCATTER", "1")})
                    r = rows[-1]
                    print(n, lay, "hints", hints, "mut", int(bool(mut)), "bits", bits, "bps", qv, r["ms"], "ms", r["Gq_per_s"],
                          "Gq/s", r["same_results"], "stages", stage, flush=True)
        ix.close()
        del ix
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"part_sweep_{n}.json"), "w"), indent=1)
