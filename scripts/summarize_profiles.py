#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag>      # e.g. r1a

Reads gpurun_out/launches_*_<tag>.csv (ncu --metrics gpu__time_duration.sum launch lists),
gpurun_out/prof_*_<tag>.ncu-rep (ncu --set full captures) and gpurun_out/bench_*_<tag>.json."""
import collections
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
os.makedirs(PR, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]

out = [f"# ncu summaries, tag {tag}\n"]
traffic = {}
layer = {}
for path in sorted(glob.glob(os.path.join(GO, f"launches_*_{tag}.csv"))):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if not rows:
        continue
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg, tot = collections.OrderedDict(), 0.0
    body = [r for r in rows[1:] if r[mi] == "gpu__time_duration.sum"]
    # keep the LAST forward only (the list also holds the weight-repacking launches and the warm-up forward)
    wl = os.path.basename(path).split("_")[1]
    logp = os.path.join(GO, f"ncu_list_{wl}.log")
    m = re.search(r"launches/forward (\d+)", open(logp).read()) if os.path.exists(logp) else None
    if m:
        body = body[-int(m.group(1)):]
    for r in body:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    out.append(f"\n## launch list {os.path.basename(path)} (cold-cache, serialised; compare shares)\n")
    out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |")
for path in sorted(glob.glob(os.path.join(GO, f"prof_*_{tag}.ncu-rep"))):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(vals, units)))
        out.append(f"\n## full capture {os.path.basename(path)}: `{d['Kernel Name'][0][:120]}`\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                out.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
        if "dram__bytes_read.sum" in d:
            def tobytes(v, u):
                m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                return float(v.replace(",", "")) * m.get(u, 1)
            tr = tobytes(*d["dram__bytes_read.sum"]) + tobytes(*d["dram__bytes_write.sum"])
            out.append(f"| dram traffic (read+write) | {tr / 1e6:.1f} | MB per launch |")
            m3 = re.match(r"prof_(qkv|attn|fc_ln|ffn_w1|w2_ln)_(c\d)_", os.path.basename(path))
            if m3:   # one decoder layer = these five kernels
                layer.setdefault(m3.group(2), {})[m3.group(1)] = (int(tr), float(d["gpu__time_duration.sum"][0]))
            m2 = re.match(r"prof_ffn_w1_(c\d)_", os.path.basename(path))
            if m2:   # bench.py's roofline.traffic reads this (dominant kernel: decoder FFN conv k=9 GEMM)
                traffic[m2.group(1)] = {"dec.ffn_w1_bytes_per_launch": int(tr), "source": os.path.basename(path),
                                        "kernel": d["Kernel Name"][0][:60], "gpu_time_us": d["gpu__time_duration.sum"][0]}
for path in sorted(glob.glob(os.path.join(GO, f"bench_*_{tag}.json"))):
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        continue
    out.append(f"\n## bench line {os.path.basename(path)}\n\n```json\n{json.dumps(d, indent=1)}\n```")
for wl, parts in layer.items():
    if len(parts) == 5:
        traffic.setdefault(wl, {})["decoder_layer_bytes"] = sum(v[0] for v in parts.values())
        traffic[wl]["decoder_layer_us_under_ncu"] = round(sum(v[1] for v in parts.values()), 1)
        out.append(f"\n## decoder layer at {wl} (qkv + attention + fc_ln + ffn_w1 + ffn_w2_ln, one capture each)\n")
        out.append("| kernel | DRAM MB | us (ncu, cold) | GB/s |\n|---|---|---|---|")
        for k, (by, us) in parts.items():
            out.append(f"| {k} | {by / 1e6:.1f} | {us:.1f} | {by / us / 1e3:.0f} |")
        tb, tu = sum(v[0] for v in parts.values()), sum(v[1] for v in parts.values())
        out.append(f"| layer | {tb / 1e6:.1f} | {tu:.1f} | {tb / tu / 1e3:.0f} |")
if traffic:
    tp = os.path.join(PR, "roofline_traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    for k, v in traffic.items():
        old.setdefault(k, {}).update(v)
    json.dump(old, open(tp, "w"), indent=1)
    print("wrote", tp, traffic)
open(os.path.join(PR, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("wrote", os.path.join(PR, f"{tag}_summary.md"))
