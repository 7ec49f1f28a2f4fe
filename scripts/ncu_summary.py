#!/usr/bin/env python
"""Summarise an ``ncu --set full`` report (one or more kernels) into the numbers the roofline line needs.

    python scripts/ncu_summary.py gpurun_out/ncu_default_step.ncu-rep --out profiles/r02/ncu_cfg2_step_summary.txt \
        [--traffic-key cfg2]      # also (re)writes that config's entry of profiles/ncu_traffic.json

Per kernel: duration, DRAM bytes read / written (-> ``roofline.traffic``), achieved DRAM GB/s, the unit
throughputs that name the limiter (L1/TEX, L2 = lts, DRAM, SM), L2 hit rate, occupancy, registers, the top
warp-stall reasons and the shared-memory bank conflicts.  Numbers taken under the profiler are cold-cache and
serialised: they explain a bench line, they are never the bench value.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}


def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    out = []
    for r in body:
        d = {}
        for name, unit, val in zip(head, units, r):
            try:
                d[name] = float(val.replace(",", "")) * UNITS.get(unit, 1)
            except ValueError:
                d[name] = val
        out.append(d)
    return out


def short(name):
    name = name.replace("kge::", "")
    return name.split("(")[0] if "<" not in name else name.split(">(")[0] + ">"


def summarise(k):
    g = k.get
    dur = g("gpu__time_duration.sum")
    rd, wr = g("dram__bytes_read.sum", 0.0), g("dram__bytes_write.sum", 0.0)
    stalls = sorted(((v, n.split("issue_stalled_")[1].split("_per_warp_active")[0].replace(".pct", ""))
                     for n, v in k.items()
                     if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_per_warp_active.pct")
                     and isinstance(v, float)), reverse=True)[:4]
    if not stalls:
        stalls = sorted(((v, n.split("issue_stalled_")[1].split(".")[0]) for n, v in k.items()
                         if "issue_stalled" in n and n.endswith(".ratio") and isinstance(v, float)), reverse=True)[:4]
    d = {
        "kernel": short(g("Kernel Name")), "grid": g("Grid Size"), "block": g("Block Size"),
        "duration_us": dur * 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr,
        "dram_gbs": (rd + wr) / dur / 1e9,
        "dram_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1tex_pct": g("l1tex__throughput.avg.pct_of_peak_sustained_active"),
        "l2_pct": g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_pct": g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
        "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "regs": g("launch__registers_per_thread"),
        "smem_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "tensor_pipe_pct": g("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                             g("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")),
        "xu_pipe_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                         g("SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed")),
        "nvlink_peer_pct": g("SYSLTS.TriageCompute.syslts__t_sector_throughput_aperture_peer.avg.pct_of_peak_sustained_elapsed"),
        "top_stalls": [f"{n}={v:.1f}" for v, n in stalls],
    }
    lim = max((("L1/TEX", d["l1tex_pct"]), ("L2", d["l2_pct"]), ("DRAM", d["dram_pct"]), ("SM", d["sm_pct"])),
              key=lambda t: t[1] if isinstance(t[1], float) else -1)
    d["limiter"] = lim[0]
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--out")
    ap.add_argument("--traffic-key")
    a = ap.parse_args()
    ks = [summarise(k) for k in load(a.rep)]
    lines = [f"# {os.path.basename(a.rep)} — ncu --set full --clock-control none; per-launch numbers (cold cache, serialised)"]
    for d in ks:
        lines.append("")
        lines.append(f"{d['kernel']}   grid {d['grid']} x block {d['block']}, {d['regs']:.0f} regs")
        lines.append(f"  duration            {d['duration_us']:9.1f} us")
        lines.append(f"  DRAM read / write   {d['dram_read_bytes'] / 1e6:9.1f} / {d['dram_write_bytes'] / 1e6:.1f} MB  "
                     f"-> {d['dram_gbs']:.0f} GB/s ({d['dram_pct']:.1f} % of DRAM peak)")
        lines.append(f"  unit throughput     L1/TEX {d['l1tex_pct']:.1f} %   L2 {d['l2_pct']:.1f} %   DRAM {d['dram_pct']:.1f} %   "
                     f"SM {d['sm_pct']:.1f} %   => limiter: {d['limiter']}")
        lines.append(f"  hit rates           L1 {d['l1_hit_pct']:.1f} %   L2 {d['l2_hit_pct']:.1f} %")
        tp = d["tensor_pipe_pct"]
        lines.append(f"  warps active        {d['warps_active_pct']:.1f} %   smem bank conflicts {d['smem_bank_conflicts']:.0f}"
                     + (f"   tensor pipe {tp:.1f} %" if isinstance(tp, float) and tp > 0 else ""))
        lines.append(f"  top stalls          {', '.join(d['top_stalls'])}")
    text = "\n".join(lines) + "\n"
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(text)
    sys.stdout.write(text)
    if a.traffic_key:
        path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        try:
            tr = json.load(open(path))
        except Exception:
            tr = {}
        entry = {"source": os.path.basename(a.rep)}
        seen = set()
        for d in ks:
            role = ("fwd" if "score_neg_kernel" in d["kernel"] else "bwd" if "score_bwd_kernel" in d["kernel"] else
                    "adam" if "adam_kernel" in d["kernel"] else "sampler" if "sample_negatives" in d["kernel"] else None)
            if role is None or role in seen:
                continue
            seen.add(role)
            entry[role] = {k: d[k] for k in ("kernel", "duration_us", "dram_read_bytes", "dram_write_bytes", "dram_pct",
                                             "l1tex_pct", "l2_pct", "sm_pct", "l2_hit_pct", "limiter")}
        entry["bwd_dram_bytes"] = int(entry["bwd"]["dram_read_bytes"] + entry["bwd"]["dram_write_bytes"]) if "bwd" in entry else None
        entry["fwd_dram_bytes"] = int(entry["fwd"]["dram_read_bytes"] + entry["fwd"]["dram_write_bytes"]) if "fwd" in entry else None
        tr[a.traffic_key] = entry
        json.dump(tr, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
