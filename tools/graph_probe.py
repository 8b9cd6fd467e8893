#!/usr/bin/env python
"""Replay time of the captured frame graph as a function of WHEN it was captured (measurement helper for DESIGN 4.5)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astre_b200 import GpuContext, gpu, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_777_216
scene = scenes.config3(n)
G = gpu.FRAME_ALL | gpu.FRAME_GRAPH


def timed(ctx, flags, steps=30):
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        ctx.frame(flags)
        torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        ctx.frame(flags)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for name, plain_before, sync_before in (("graph from the first frame", 0, False), ("3 plain frames, sync, then graph", 3, True),
                                        ("3 plain frames, no sync, then graph", 3, False), ("plain only", -1, False)):
    with GpuContext(scene.n, max_views=5) as ctx:
        scenes.load_into(ctx, scene)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            ctx.set_stream(stream.cuda_stream)
            for _ in range(max(plain_before, 0)):
                ctx.frame(gpu.FRAME_ALL)
            if sync_before:
                ctx.sync()
                time.sleep(0.05)
            ms = timed(ctx, gpu.FRAME_ALL if plain_before < 0 else G)
            ms2 = timed(ctx, gpu.FRAME_ALL if plain_before < 0 else G)
            print(f"{name:40s} {ms:.4f} then {ms2:.4f} ms/frame   graph (captures, view patches, bucket bits) = {ctx.graph_stats()}", flush=True)
            ctx.set_stream(None)
