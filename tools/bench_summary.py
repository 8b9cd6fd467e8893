#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line (headline + extra workloads).  Usage: bench_summary.py file.json ..."""
import json
import sys


def show(name, w):
    r, e = w["roofline"], w["e2e"]
    print(f"{name}: {w['ms_per_step']:.4f} ms/step -> {w['value'] / 1e9:.2f} G entities/s | kernel {r['kernel_ms']:.4f} ms, "
          f"{r['achieved']:.0f} GB/s = {r['frac']:.3f} of peak | launches {w['gpu_launches']} sort {w.get('sort_path')} graph {w.get('frame_as_cuda_graph')}")
    print(f"    e2e {e['ms_per_step']:.3f} ms -> {e['value'] / 1e6:.1f} M/s | dirty 1% {e['e2e_dirty_1pct']['ms_per_step']:.3f} ms "
          f"{e['e2e_dirty_1pct']['value'] / 1e6:.0f} M/s | dirty 10% {e['e2e_dirty_10pct']['ms_per_step']:.3f} ms {e['e2e_dirty_10pct']['value'] / 1e6:.0f} M/s")
    c, p = w.get("cpu_baseline"), w["parity"]
    if c:
        print(f"    cpu {c['value'] / 1e6:.1f} M/s ({c['cores']} threads)", end="")
    print(f"    parity {p['ranks_ok']}/{p['ranks']} {p['rank0']}")


for path in sys.argv[1:]:
    line = [l for l in open(path) if l.startswith("{")][-1]
    d = json.loads(line)
    print(f"== {path}: n_gpus {d['n_gpus']} scaling {d['scaling']}")
    head = dict(d)
    head["sort_path"] = d.get("sort_path", d["config"].get("sort_path"))
    head["frame_as_cuda_graph"] = d.get("frame_as_cuda_graph", d["config"].get("frame_as_cuda_graph"))
    show("headline " + d["config"]["workload"][:40], head)
    for k, v in (d.get("extra_workloads") or d["config"].get("extra_workloads", {})).items():
        if "error" in v:
            print(k, v)
        else:
            show(k, v)
    print("   clocks", d.get("clocks"))
