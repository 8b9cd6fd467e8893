#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the transform + visibility path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): entities transformed+culled per second at 16M entities.
A "step" is one frame of the hot path (fused transform + 5-view frustum cull +
compaction + draw-key radix sort) over one resident batch of 16,777,216 entities
per GPU.  N=1 runs configuration 3; N>1 gives every rank its own 16M-entity range
of the configuration-4 stream (weak scaling) and adds the per-frame NCCL
all-gather of visible counts.

`value` is device-resident throughput (CUDA events, max over ranks); `e2e` pushes
every entity's TRS from pinned host memory, runs the frame and reads back counts,
sorted keys and the visible world matrices, all through the C ABI, inside the
timed region.  `--impl reference` times the CPU oracle (the port of the
reference's algorithm -- the reference itself cannot be built here)
on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 16_777_216
METRIC = "entities transformed+culled/sec at 16M entities"
BYTES_PER_ENTITY = 108        # 40 B TRS + 4 B meta read, 64 B world matrix written
BYTES_PER_KEY = 8


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("k_transform_cull_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(rank, world):
    from astre_b200 import scenes
    if world == 1:
        return scenes.config3(N_PER_GPU), "config3: 16,777,216 flat entities, camera + 4 shadow-cascade frusta"
    return (scenes.config4(N_PER_GPU * world, rank * N_PER_GPU, (rank + 1) * N_PER_GPU),
            f"config4 stream: {world} x 16,777,216 flat entities sharded by entity range, camera + 4 cascades")


def cpu_port_frame(scene_arrays, n, threads=None):
    """One oracle frame (transform + cull + sort) over the first n entities; returns seconds."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import oracle_lib
    o = oracle_lib.load()
    if threads:
        o.lib.orc_set_threads(int(threads))
    pos, rot, scale, mesh, shader, flags, bounds, clip, masks = scene_arrays
    brecs, nb = o.make_bounds(bounds)
    vrecs = o.make_views(clip, masks)
    F = len(masks)
    world = np.empty((n, 16), np.float32)
    cap = n * F
    keys = np.empty(cap, np.uint64)
    counts = np.zeros(F, np.uint32)
    p = oracle_lib._p
    t0 = time.perf_counter()
    total = o.lib.orc_frame(C.c_uint32(n), p(pos, C.c_float), p(rot, C.c_float), p(scale, C.c_float), None,
                            p(mesh, C.c_uint32), p(shader, C.c_uint32), p(flags, C.c_uint8), brecs, C.c_uint32(nb),
                            C.c_uint32(1), C.c_uint32(0), C.c_uint32(F), vrecs, p(world, C.c_float),
                            p(keys, C.c_uint64), C.c_uint64(cap), p(counts, C.c_uint32))
    dt = time.perf_counter() - t0
    return dt, int(total), o.lib.orc_max_threads()


def cpu_reference_layout(scene, n):
    """B0 of SURVEY.md 8d: TransformSystem::run + VisualSystem::run over the reference's storage shape (hash map of
    per-entity structs, one thread -- oracle/ref_store.cpp).  Returns entities/s."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    so = os.path.join(oracle_lib.ORACLE_DIR, "_build", "libastre_refstore.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", oracle_lib.ORACLE_DIR], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.refstore_create.restype = C.c_void_p
    lib.refstore_visual.restype = C.c_uint64
    p = oracle_lib._p
    st = lib.refstore_create(C.c_uint32(n), p(scene.pos, C.c_float), p(scene.rot, C.c_float), p(scene.scale, C.c_float),
                             p(scene.mesh, C.c_uint32), p(scene.shader, C.c_uint32), p(scene.flags, C.c_uint8))
    try:
        t0 = time.perf_counter()
        lib.refstore_transform(C.c_void_p(st))
        lib.refstore_visual(C.c_void_p(st))
        return n / (time.perf_counter() - t0)
    finally:
        lib.refstore_destroy(C.c_void_p(st))


def scene_arrays(scene, n):
    return (scene.pos[:n], scene.rot[:n], scene.scale[:n], scene.mesh[:n], scene.shader[:n], scene.flags[:n],
            scene.bounds, scene.clip, scene.masks)


def run_reference(args):
    """CPU arm: the oracle port on all host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from astre_b200 import scenes
    n = 4_194_304
    scene = scenes.config3(n)
    arrs = scene_arrays(scene, n)
    for _ in range(max(args.warmup, 1)):
        cpu_port_frame(arrs, n)
    times = []
    for _ in range(args.steps):
        dt, total, threads = cpu_port_frame(arrs, n)
        times.append(dt)
    sec = float(np.mean(times))
    value = n / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "entities/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config3: flat entities, camera + 4 shadow-cascade frusta; transform + cull + key sort",
                   "entities_per_step": n},
        "cpu_baseline": {"value": value, "unit": "entities/s", "cores": threads, "kind": "port",
                         "sample": f"first {n} entities of config3 per step (oracle/astre_oracle.c, OpenMP, all host cores); "
                                   "the reference itself cannot be built in this image"},
        "e2e": {"value": value, "unit": "entities/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="astre_b200")
    ap.add_argument("--entities", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from astre_b200 import GpuContext, gpu, scenes

    if args.entities:
        globals()["N_PER_GPU"] = args.entities
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: astre_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene, workload = build_scene(rank, world)
    F = scene.views_per_world
    n = scene.n
    ctx = GpuContext(n, max_views=F, device=local_rank)
    scenes.load_into(ctx, scene)
    # a real (non-default) stream: the library enqueues everything on it and the
    # torch events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    from astre_b200 import multi

    class _DeviceCounts:      # zero-copy torch view of the library's per-view counters
        __cuda_array_interface__ = {"shape": (F,), "typestr": "<i4", "data": (ctx.device_counts_ptr(), False), "version": 2}
    counts_dev = torch.as_tensor(_DeviceCounts(), device="cuda")

    side = torch.cuda.Stream() if world > 1 else None
    cull_done = torch.cuda.Event() if world > 1 else None

    def step():
        if world == 1:
            ctx.frame(gpu.FRAME_ALL)
            return
        # The per-view counts are final when the cull kernel ends: the only exchange on the path (NCCL all-gather of
        # the per-shard counts) runs on a side stream while the main stream sorts this shard's keys.
        ctx.frame(gpu.FRAME_TRANSFORM | gpu.FRAME_CULL | gpu.FRAME_SORT_LATER)
        cull_done.record(stream)
        side.wait_event(cull_done)
        with torch.cuda.stream(side):
            multi.exchange_counts(counts_dev)
        ctx.sort()
        stream.wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()     # clocks are sampled from warm-up to the end of the kernel-timing loop
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    ms_step = float(t.item()) / args.steps
    value = n * world / (ms_step * 1e-3)

    # dominant kernel (fused transform+cull+compact), timed by the library's own CUDA events
    cull_ms = []
    for _ in range(args.steps):
        ctx.frame(gpu.FRAME_ALL)
        cull_ms.append(ctx.last_cull_ms())
    total_keys = ctx.read_total()
    counts = ctx.read_counts()
    # keep the GPU under the same load for a moment so that the 100 ms sampler sees it
    t_end = time.perf_counter() + 1.0
    while time.perf_counter() < t_end:
        for _ in range(50):
            ctx.frame(gpu.FRAME_ALL)
        ctx.sync()
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms = float(np.mean(cull_ms))
    alg_bytes = n * BYTES_PER_ENTITY + total_keys * BYTES_PER_KEY
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": profiled_traffic(), "kernel": "k_frame_fused<false,false,true>", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": kernel_ms / ms_step}

    # end to end through the C ABI with host buffers (pinned), copies inside the timed region
    pos_h = torch.from_numpy(scene.pos).pin_memory()
    rot_h = torch.from_numpy(scene.rot).pin_memory()
    scl_h = torch.from_numpy(scene.scale).pin_memory()
    keys_h = torch.empty(max(total_keys, 1), dtype=torch.int64).pin_memory()
    vis_h = torch.empty((max(total_keys, 1), 16), dtype=torch.float32).pin_memory()
    pos_np, rot_np, scl_np = pos_h.numpy(), rot_h.numpy(), scl_h.numpy()
    keys_np, vis_np = keys_h.numpy().view(np.uint64), vis_h.numpy()
    cnt_np = np.zeros(F, np.uint32)

    def e2e_step():
        ctx.update_trs_range(0, n, pos_np, rot_np, scl_np)
        ctx.frame(gpu.FRAME_ALL)
        ctx.read_counts(cnt_np)
        k = ctx.read_all_keys(keys_np)
        ctx.read_visible_world(0, len(k), vis_np)
        return len(k)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    e0.record(stream)
    for _ in range(e2e_steps):
        nk = e2e_step()
    e1.record(stream)
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / e2e_steps
    e2e = {"value": n * world / (e2e_ms * 1e-3), "unit": "entities/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(F * 4 + nk * 8 + nk * 64),
           "what": "astre_gpu_update_trs_range(all entities, pinned host) + astre_gpu_frame + read_counts + "
                   "read_all_keys + read_visible_world"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        reps, dts = 3, []
        arrs = scene_arrays(scene, n)
        for _ in range(reps):
            dt, tot, threads = cpu_port_frame(arrs, n)
            dts.append(dt)
        cpu = {"value": n / float(np.mean(dts)), "unit": "entities/s", "cores": threads, "kind": "port",
               "sample": f"{reps} full frames of the same {n}-entity workload (oracle/astre_oracle.c, OpenMP, all host cores)",
               "keys_match_gpu": bool(tot == total_keys),
               "reference_layout_1thread": {"value": cpu_reference_layout(scene, 1_048_576), "unit": "entities/s",
                                            "what": "transform + proxy emission over a hash map of per-entity structs, 1 thread, "
                                                    "first 1,048,576 entities (oracle/ref_store.cpp; no culling, no sort)"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "entities/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload + "; fused transform+cull+compact + 64-bit key radix sort"
                                   + ("; NCCL all-gather of per-shard visible counts" if world > 1 else ""),
                       "entities_per_gpu": n, "views": F, "visible_keys_rank0": int(total_keys),
                       "counts_rank0": [int(c) for c in counts],
                       "l2": f"inputs ({n * 44 / 1e6:.0f} MB/GPU) and outputs ({n * 64 / 1e6:.0f} MB/GPU) exceed the 126 MB L2; no flush"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(ln.item()), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
