#!/usr/bin/env python
"""Small frames of every kernel path for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astre_b200 import GpuContext, gpu, scenes  # noqa: E402

for name, scene in (("flat", scenes.config3(6000)), ("hier", scenes.config2(n=9000)),
                    ("worlds", scenes.config5(n_worlds=4, entities_per_world=1024))):
    with GpuContext(scene.n, max_views=max(scene.views_per_world, 1)) as ctx:
        scenes.load_into(ctx, scene)
        ctx.frame(gpu.FRAME_ALL | gpu.FRAME_BASIS)
        keys = ctx.read_all_keys()
        ctx.read_runs()
        ctx.snapshot_trs()
        ctx.interpolate(0.5)
        ctx.read_visible_world(0, min(len(keys), 100))
        print(name, scene.n, len(keys), flush=True)
# dense: every entity in every view (dense cull path, multi-tile sort)
s = scenes.config1(5000)
s.pos = (s.pos * np.float32(0.02)).astype(np.float32)
s.pos[:, 2] -= np.float32(40.0)
with GpuContext(s.n, max_views=1) as ctx:
    scenes.load_into(ctx, s)
    ctx.frame()
    print("dense", len(ctx.read_all_keys()), flush=True)
