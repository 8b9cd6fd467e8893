#!/usr/bin/env python
"""Frame timing of the BASELINE.json configurations that are not the bench line (device-resident,
library CUDA events).  Usage: config_bench.py [config2|config4|config5|dense] ..."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astre_b200 import GpuContext, gpu, scenes  # noqa: E402


def run(name, scene, bytes_per_entity):
    with GpuContext(scene.n, max_views=max(scene.views_per_world, 1)) as ctx:
        scenes.load_into(ctx, scene)
        for _ in range(3):
            ctx.frame()
        cull, frame = [], []
        for _ in range(15):
            ctx.frame()
            cull.append(ctx.last_cull_ms())
            frame.append(ctx.last_frame_ms())
        tot = ctx.read_total()
        b = scene.n * bytes_per_entity + tot * 8
        print(f"{name}: n={scene.n} keys={tot} transform+cull {np.mean(cull):.4f} ms ({b / np.mean(cull) / 1e6:.0f} GB/s) "
              f"frame {np.mean(frame):.4f} ms -> {scene.n / np.mean(frame) / 1e6:.2f} G entities/s", flush=True)


for which in (sys.argv[1:] or ["config2", "config5", "config4", "dense"]):
    t0 = time.time()
    if which == "config1":
        run("config1 (10k flat, 1 view)", scenes.config1(), 108)
    elif which.startswith("flat"):
        run(f"config3 stream, {which[4:]} entities, 5 views", scenes.config3(int(which[4:])), 108)
    elif which == "config2":
        run("config2 (1M, 8-level hierarchy, 1 view)", scenes.config2(), 108 + 68 * 254 / 255)
    elif which == "config5":
        run("config5 (4096 worlds x 16384, 1 view each)", scenes.config5(), 108)
    elif which == "config4":
        run("config4 (64M flat, 5 views, 1 GPU)", scenes.config4(), 108)
    elif which == "dense":
        s = scenes.config1(4_194_304)
        s.pos = (s.pos * np.float32(0.05)).astype(np.float32)
        s.pos[:, 2] -= np.float32(40.0)
        run("dense (4M entities, ~all visible, 1 view)", s, 108)
    print(f"   (setup+run {time.time() - t0:.0f} s)", flush=True)
