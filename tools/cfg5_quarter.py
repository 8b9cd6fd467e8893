import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from astre_b200 import GpuContext, gpu, scenes
s = scenes.config5(n_worlds=1024)
with GpuContext(s.n, max_views=1) as ctx:
    scenes.load_into(ctx, s)
    for _ in range(3): ctx.frame()
    c = []
    for _ in range(10):
        ctx.frame(); c.append(ctx.last_cull_ms())
    print("config5/4:", s.n, ctx.read_total(), np.mean(c), s.n * 108 / np.mean(c) / 1e6)
