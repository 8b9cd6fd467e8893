#!/usr/bin/env python
"""Times the dominant kernel (and the whole frame) on configuration 3 with the library's own
CUDA events.  Usage: kernel_bench.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astre_b200 import GpuContext, gpu, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_777_216
scene = scenes.config3(n)
ref_keys = None
for var in [0]:
    with GpuContext(scene.n, max_views=scene.views_per_world) as ctx:
        scenes.load_into(ctx, scene)
        for _ in range(3):
            ctx.frame()
        cull, frame = [], []
        for _ in range(20):
            ctx.frame()
            cull.append(ctx.last_cull_ms())
            frame.append(ctx.last_frame_ms())
        keys = ctx.read_all_keys()
        tot = ctx.read_total()
        if ref_keys is None:
            ref_keys = keys.copy()
        same = np.array_equal(keys, ref_keys)
        b = n * 108 + tot * 8
        print(f"frame: transform+cull {np.mean(cull):.4f} ms (min {np.min(cull):.4f})  frame {np.mean(frame):.4f} ms  "
              f"{b / np.mean(cull) / 1e6:.0f} GB/s  keys {tot} same_as_first={same}", flush=True)
        # transform only
        t = []
        for _ in range(10):
            ctx.frame(gpu.FRAME_TRANSFORM)
            t.append(ctx.last_cull_ms())
        print(f"transform only: {np.mean(t):.4f} ms  {n * 104 / np.mean(t) / 1e6:.0f} GB/s", flush=True)
