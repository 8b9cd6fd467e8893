#!/usr/bin/env python
"""How much do -0 cases (axis-aligned rotations with negative components, mirrored scales) cost?  They take the
out-of-line exact chain in trs_matrix."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astre_b200 import GpuContext, gpu, scenes

n = 16_777_216
scene = scenes.config3(n)
for frac in (0.0, 0.01, 0.1, 0.5, 1.0):
    s = scene
    rot = scene.rot.copy()
    k = int(n * frac)
    if k:
        idx = np.arange(0, n, max(1, n // k))[:k]
        rot[idx] = np.array([0.0, -0.70710678, 0.0, 0.70710678], np.float32)
    with GpuContext(n, max_views=5) as ctx:
        ctx.upload_entities(0, s.pos, rot, s.scale, None, s.mesh, s.shader, s.flags)
        ctx.set_mesh_bounds(s.bounds)
        ctx.set_views(s.clip[0], s.masks)
        for _ in range(3):
            ctx.frame()
        t = []
        for _ in range(10):
            ctx.frame()
            t.append(ctx.last_cull_ms())
        print(f"fraction of entities on the exact chain {frac:5.2f}: transform+cull {np.mean(t):.4f} ms", flush=True)
