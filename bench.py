#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the transform + visibility path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1..5] [--no-extra]

Metric (BASELINE.json): entities transformed+culled per second at 16M entities.  A "step" is one frame of the
hot path (fused transform + frustum cull + compaction + 64-bit draw-key sort) over one resident batch.

  headline   N = 1: configuration 3 (16,777,216 flat entities, camera + 4 shadow-cascade frusta).
             N > 1: every rank owns its own 16,777,216-entity range of the configuration-4 stream (weak scaling)
             and the per-frame exchange of SURVEY.md 8e runs every step, on a side stream, off the frame's critical
             path: one kernel stores the per-shard visible counts into every rank's count board over NVLink, waits
             for everybody's and derives this rank's segment offsets (astre_gpu_count_board_exchange; no collective
             library on the path); --exchange nccl uses an NCCL all-gather + astre_gpu_global_offsets_dev instead.
             The timed region ends only when the last exchange has finished.
  extras     (top-level key `extra_workloads`, unless --no-extra) the other BASELINE.json configurations, each with its own
             value / roofline / e2e / cpu_baseline / parity: config 1 (10k flat, N = 1 only), config 2 (1M, 8-level
             hierarchy; root-subtree shards), config 4 (67,108,864 flat FIXED, entity-range shards: strong
             scaling), config 5 (4096 worlds x 16,384 FIXED, whole-world shards: strong scaling).
  --config C makes configuration C the headline workload instead (own line, no extras).

`value` is device-resident throughput (CUDA events on the launching stream, max over ranks); `e2e` pushes every
entity's TRS from pinned host memory, runs the frame and reads back counts, sorted keys and the visible world
matrices, all through the C ABI, inside the timed region (plus e2e_dirty_1pct / _10pct: the slot-list update the
mirror was built for).  After the timed regions every rank replays its shard with the CPU oracle and compares
counts, the full sorted key list and every world matrix bit for bit (`parity`).

`--impl reference` times the CPU side of the comparison on the SAME workload and size (its `config` object equals
the GPU arm's), on all host cores (threads set explicitly): the oracle port of the path (transform + cull + sort;
the reference itself has no culling or sort, SURVEY.md section 0, so its own sources cannot run the whole path),
plus, as a sub-figure, the reference's OWN TransformSystem::run + VisualSystem::run compiled from its sources
(oracle/_ref) on its own hash-map storage, one thread.  That process never loads the product library.
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
BYTES_PER_CHILD = 68          # + parent index 4 B + parent matrix 64 B for a non-root entity
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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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


# ---- workloads ------------------------------------------------------------------------------------------------

def make_workload(name, rank, world):
    """-> (scene of this rank, description, scaling, entities over all ranks)"""
    from astre_b200 import scenes
    if name == "config1":
        return scenes.config1(), "config1: 10,000 flat entities, single camera", "replicas", 10_000 * world
    if name == "config2":
        s = scenes.config2()
        total = s.n
        if world > 1:
            s = s.shard(rank, world)
        return s, "config2: 1,048,576 entities, 8-level hierarchy, single camera" + \
            (f"; sharded by root subtree over {world} GPUs" if world > 1 else ""), "strong", total
    if name == "config3":
        return scenes.config3(N_PER_GPU), f"config3: {N_PER_GPU:,} flat entities, camera + 4 shadow-cascade frusta", "weak", N_PER_GPU * world
    if name == "config4w":
        return (scenes.config4(N_PER_GPU * world, rank * N_PER_GPU, (rank + 1) * N_PER_GPU),
                f"config4 stream: {world} x {N_PER_GPU:,} flat entities sharded by entity range, camera + 4 cascades", "weak",
                N_PER_GPU * world)
    if name == "config4":
        total = 67_108_864
        per = total // world
        return (scenes.config4(total, rank * per, (rank + 1) * per),
                f"config4: 67,108,864 flat entities (fixed), entity-range shards over {world} GPU(s), camera + 4 cascades",
                "strong", total)
    if name == "config5":
        W, E = 4096, 16_384
        per = W // world
        return (scenes.config5(W, E, rank * per, (rank + 1) * per),
                f"config5: 4096 independent worlds x 16,384 entities (fixed), whole-world shards over {world} GPU(s), one camera per world",
                "strong", W * E)
    raise SystemExit(f"unknown workload {name}")


def headline_config(workload, n, views, visible_keys, world):
    """The `config` object of the JSON line -- built by BOTH arms from the same values, so that they compare equal."""
    return {"workload": workload + "; transform + frustum cull + compaction + 64-bit draw-key sort"
                        + ("; per-frame exchange of per-shard visible counts and segment offsets" if world > 1 else ""),
            "entities_per_gpu": int(n), "views": int(views), "visible_keys_rank0": int(visible_keys),
            "l2": f"inputs ({n * 44 / 1e6:.0f} MB/GPU) and outputs ({n * 64 / 1e6:.0f} MB/GPU) exceed the 126 MB L2; no flush"
                  if n * 108 > 4 * 126e6 else "working set fits the 126 MB L2: launch-latency regime, stated as such"}


def algorithmic_bytes(scene, total_keys):
    """HBM bytes one launch of the transform + cull kernel must move."""
    children = 0 if scene.parent is None else int((scene.parent != 0xFFFFFFFF).sum())
    return scene.n * BYTES_PER_ENTITY + children * BYTES_PER_CHILD + total_keys * BYTES_PER_KEY


# ---- CPU side (oracle port; the only place this file executes oracle/) ---------------------------------------------

def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib


def cpu_port_frame(scene, threads):
    """One oracle frame (transform + cull + sort) of a scene -> (seconds, world, keys, counts, threads used)."""
    import ctypes as C
    ol = _oracle()
    o = ol.load()
    o.lib.orc_set_threads(int(threads))
    n, F, W = scene.n, scene.views_per_world, scene.n_worlds
    brecs, nb = o.make_bounds(scene.bounds)
    vrecs = o.make_views(scene.clip, scene.masks)
    world = np.empty((n, 16), np.float32)
    cap = max(n * F, 1)
    keys = np.empty(cap, np.uint64)
    counts = np.zeros(max(W * F, 1), np.uint32)
    p = ol._p
    parent = None if scene.parent is None else np.ascontiguousarray(scene.parent, np.uint32)
    o.lib.orc_frame.restype = C.c_uint64
    t0 = time.perf_counter()
    total = o.lib.orc_frame(C.c_uint32(n), p(scene.pos, C.c_float), p(scene.rot, C.c_float), p(scene.scale, C.c_float),
                            p(parent, C.c_uint32), p(scene.mesh, C.c_uint32), p(scene.shader, C.c_uint32), p(scene.flags, C.c_uint8),
                            brecs, C.c_uint32(nb), C.c_uint32(W), C.c_uint32(scene.entities_per_world), C.c_uint32(F), vrecs,
                            p(world, C.c_float), p(keys, C.c_uint64), C.c_uint64(cap), p(counts, C.c_uint32))
    dt = time.perf_counter() - t0
    return dt, world, keys[:int(total)], counts[:W * F], int(o.lib.orc_max_threads())


def reference_systems_tick(scene, n):
    """The reference's OWN TransformSystem::run + VisualSystem::run (oracle/_ref, compiled from its sources) over its
    own storage (hash maps of component messages), one thread as the reference runs them -> entities/s, or None."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import ref_lib
        if not os.path.exists(ref_lib.SO):
            return None
        n = min(n, scene.n)
        ids = np.arange(1, n + 1, dtype=np.uint64)
        with ref_lib.RefWorld() as w:
            w.register_vertex_buffer("mesh", 1)
            w.register_shader("shader", 1)
            w.add_transforms(ids, scene.pos[:n], scene.rot[:n], scene.scale[:n])
            for e in ids:
                w.add_visual(int(e), True, "mesh", "shader")
            w.time_tick(1, True)
            return n / w.time_tick(1, True)
    except Exception:
        return None


class OracleMatrices:
    """Clip matrices of the synthetic views built by the oracle's GLM restatement (identical bits to the product's
    host math, tests/test_host_math.py) -- so that the reference arm never maps the product library."""

    @staticmethod
    def camera_clip(pos, fwd, up, fov_deg, aspect, z_near, z_far):
        o = _oracle().load()
        view, proj = o.camera_view_proj(pos, fwd, up, fov_deg, aspect, z_near, z_far)
        return o.mat4_mul(proj, view)

    @staticmethod
    def ortho_view(eye, center, up, l, r, b, t, n, f):
        o = _oracle().load()
        return o.mat4_mul(o.ortho(l, r, b, t, n, f), o.look_at(eye, center, up))


def run_reference(args):
    """CPU arm: same workload, same entity count as the GPU arm's rank 0, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import astre_b200._lib as product
    from astre_b200 import scenes
    scenes.set_matrix_provider(OracleMatrices)
    name = f"config{args.config}" if args.config else ("config3" if world == 1 else "config4w")
    scene, workload, scaling, _ = make_workload(name, 0, world)
    threads = host_threads()
    for _ in range(max(args.warmup, 1)):
        cpu_port_frame(scene, threads)
    times = []
    for _ in range(args.steps):
        dt, _, keys, _, used = cpu_port_frame(scene, threads)
        times.append(dt)
    sec = float(np.mean(times))
    value = scene.n / sec
    tick = reference_systems_tick(scene, 262_144)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "entities/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(workload, scene.n, scene.views_per_world, len(keys), world),
        "cpu_baseline": {"value": value, "unit": "entities/s", "cores": used, "kind": "port",
                         "sample": f"all {scene.n} entities of rank 0's workload per step (oracle/astre_oracle.c, OpenMP, "
                                   f"{used} threads set explicitly)",
                         "why_port": "the path of the metric includes frustum culling, compaction and the key sort, which the reference "
                                     "does not have (SURVEY.md section 0): the oracle port is the only CPU implementation of the whole "
                                     "path; the part the reference does implement is timed from its own sources below",
                         "reference_systems_1thread": None if tick is None else {
                             "value": tick, "unit": "entities/s", "kind": "reference",
                             "what": "the reference's own TransformSystem::run + VisualSystem::run (oracle/_ref, compiled from "
                                     "/root/reference sources) on its hash-map storage, first 262,144 entities, 1 thread; "
                                     "protobuf messages are stand-in structs, so this flatters the reference"}},
        "e2e": {"value": value, "unit": "entities/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "product_library_loaded": product._LIB is not None,
    }
    print(json.dumps(line), flush=True)


# ---- GPU side -------------------------------------------------------------------------------------------------

class Bench:
    """One workload on this rank's GPU: device-resident timing, dominant-kernel timing, end to end, parity."""

    def __init__(self, name, rank, local_rank, world, torch, dist, use_graph=True, exchange="p2p"):
        from astre_b200 import GpuContext, gpu, multi, scenes
        self.torch, self.dist, self.gpu, self.multi = torch, dist, gpu, multi
        self.name, self.rank, self.world = name, rank, world
        self.scene, self.workload, self.scaling, self.total_entities = make_workload(name, rank, world)
        s = self.scene
        self.F, self.n = s.views_per_world, s.n
        self.n_counts = s.n_worlds * s.views_per_world
        self.ctx = GpuContext(max(s.n, 1), max_views=self.F, device=local_rank)
        scenes.load_into(self.ctx, s)
        # a real (non-default) stream: the library enqueues everything on it and the torch events are recorded on it
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.ctx.set_stream(self.stream.cuda_stream)
        self.rank_major = s.n_worlds > 1          # whole-world shards: the global list is the concatenation of the ranks' lists
        self.exchange_kind = exchange
        self.exchange = None
        if (world > 1 or os.environ.get("ASTRE_BENCH_FORCE_EXCHANGE")) and name != "config1":     # (the variable: exchange path on one GPU)
            if exchange == "p2p":
                # count boards need CUDA IPC between the ranks' processes; if any rank cannot map its peers (a container
                # without shared PID / IPC namespaces), ALL ranks use the NCCL variant and the line says so
                err = ""
                try:
                    self.exchange = multi.PeerCountExchange(self.ctx, self.n_counts, rank_major=self.rank_major)
                except Exception as e:      # noqa: BLE001
                    err = f"{type(e).__name__}: {e}"
                if self.max_over_ranks(1.0 if err else 0.0) > 0:
                    print(f"bench.py: rank {rank}: count boards unavailable ({err or 'on another rank'}); using the NCCL exchange",
                          file=sys.stderr, flush=True)
                    self.exchange = None
                    self.exchange_kind = "nccl (count boards unavailable)"
            if self.exchange is None:
                self.exchange = multi.CountExchange(self.ctx, self.n_counts, rank_major=self.rank_major)
        # steps replay the frame as one CUDA graph (ASTRE_FRAME_GRAPH); --no-graph launches kernel by kernel
        self.frame_flags = gpu.FRAME_ALL | (gpu.FRAME_GRAPH if use_graph else 0)
        self.e2e_flags = self.frame_flags | gpu.FRAME_GATHER     # end-to-end steps read the list's matrices back: gathered in-frame

    def close(self):
        self.torch.cuda.synchronize()
        if self.exchange:
            self.exchange.close()
        self.ctx.close()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def step(self):
        if self.exchange is None:
            self.ctx.frame(self.frame_flags)
            return
        # the frame's last kernel mirrors the counts into a slot; the exchange (count boards: one kernel over peer memory;
        # nccl: all-gather + offsets kernel) runs on a side stream while the next frames compute
        self.exchange.before_frame()
        self.ctx.frame(self.frame_flags)
        self.exchange.after_frame()

    def max_over_ranks(self, x, op="max"):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def device_resident(self, steps, warmup):
        torch = self.torch
        for _ in range(max(warmup, 3)):
            self.step()
        if self.exchange:
            self.exchange.finish()
        self.barrier()
        # The frame graph is captured once a finished frame has told the host how many keys to size the sort's buckets for
        # (astre_gpu.cu: plan_settled).  Frames enqueued back to back before that ran kernel by kernel; four more untimed
        # steps, each followed by a synchronisation, make sure the capture does not land inside the timed region.
        for _ in range(4):                 # (the bucket width also settles over the first finished frames: one bit per frame)
            self.step()
            if self.exchange:
                self.exchange.finish()
            self.barrier()
        launches0 = self.ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for _ in range(steps):
            self.step()
        if self.exchange:
            self.exchange.finish()          # the last exchanges end inside the timed region
        e1.record(self.stream)
        self.barrier()
        own_ms = e0.elapsed_time(e1) / steps
        ms_step = self.max_over_ranks(own_ms)
        launches = int(self.max_over_ranks(self.ctx.launch_count - launches0, "sum"))
        self.scaling_detail = None
        if self.exchange:
            # what the exchange and the slowest GPU cost: the same loop without the exchange, and the spread over ranks
            self.barrier()
            e0.record(self.stream)
            for _ in range(steps):
                self.ctx.frame(self.frame_flags)
            e1.record(self.stream)
            self.barrier()
            plain = e0.elapsed_time(e1) / steps
            self.scaling_detail = {"ms_per_step_min_over_ranks": -self.max_over_ranks(-own_ms), "ms_per_step_max_over_ranks": ms_step,
                                   "without_exchange_ms_min_over_ranks": -self.max_over_ranks(-plain),
                                   "without_exchange_ms_max_over_ranks": self.max_over_ranks(plain),
                                   "exchange": self.exchange_kind,
                                   "note": "ranks are coupled through the exchange (a count slot is reused only when every rank has consumed "
                                           "it), so all ranks run at the pace of the slowest GPU; "
                                           "exchange cost = max - without_exchange max"}
        # torch-issued kernels inside the step at N > 1: staging copy + NCCL all-gather per frame (not counted as ours)
        return ms_step, launches

    def kernel_timing(self, steps):
        """Dominant kernel (transform + cull + compaction + keys), timed by the library's own CUDA events."""
        ms = []
        for _ in range(steps):
            self.ctx.frame(self.gpu.FRAME_ALL)
            ms.append(self.ctx.last_cull_ms())
        self.total_keys = self.ctx.read_total()
        self.counts = self.ctx.read_counts()
        self.sort_path = self.ctx.last_sort_path()
        return float(np.mean(ms))

    def roofline(self, kernel_ms, ms_step):
        alg = algorithmic_bytes(self.scene, self.total_keys)
        peak, peak_src = measured_peak()
        achieved = alg / (kernel_ms * 1e-3) / 1e9
        kern = {"config2": "k_frame_fused<true,false,true> (+ k_transform_cull_indexed for small levels)",
                "config5": "k_frame_fused<false,true,true>"}.get(self.name, "k_frame_fused<false,false,true>")
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": profiled_traffic() if self.name == "config3" else None, "kernel": kern, "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg, "peak_source": peak_src, "kernel_share_of_step": kernel_ms / ms_step}

    def end_to_end(self, steps):
        """Through the C ABI with host buffers (pinned), copies inside the timed region."""
        torch, s, g, ctx, n = self.torch, self.scene, self.gpu, self.ctx, self.n
        pos_h, rot_h, scl_h = (torch.from_numpy(a).pin_memory() for a in (s.pos, s.rot, s.scale))
        nk0 = max(self.total_keys, 1)
        keys_h = torch.empty(nk0 + nk0 // 8 + 1024, dtype=torch.int64).pin_memory()
        vis_h = torch.empty((len(keys_h), 16), dtype=torch.float32).pin_memory()
        pos_np, rot_np, scl_np = pos_h.numpy(), rot_h.numpy(), scl_h.numpy()
        keys_np, vis_np = keys_h.numpy().view(np.uint64), vis_h.numpy()
        cnt_np = np.zeros(max(self.n_counts, 1), np.uint32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if n <= 2_000_000:
            steps = max(steps, 20)          # sub-millisecond steps: three of them would be all noise

        def timed(fn, reps):
            for _ in range(3):
                fn()
            self.barrier()
            e0.record(self.stream)
            for _ in range(reps):
                nk = fn()
            e1.record(self.stream)
            self.barrier()
            return self.max_over_ranks(e0.elapsed_time(e1)) / reps, nk

        def read_results():
            ctx.read_counts(cnt_np)
            return ctx.read_draw_list(keys_np, vis_np)       # sorted keys + the visible entities' world matrices

        def set_views():           # the engine hands the frame's view / light-space matrices over every tick (visual_system glue)
            if s.n_worlds > 1:
                ctx.set_world_views(s.n_worlds, s.entities_per_world, s.clip, s.masks)
            else:
                ctx.set_views(s.clip[0], s.masks)

        def all_dirty():
            ctx.update_trs_range(0, n, pos_np, rot_np, scl_np)
            set_views()
            ctx.frame(self.e2e_flags)
            return read_results()

        ms, nk = timed(all_dirty, steps)
        out = {"value": self.total_entities / (ms * 1e-3), "unit": "entities/s", "ms_per_step": ms,
               "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(self.n_counts * 4 + nk * 8 + nk * 64),
               "what": "astre_gpu_update_trs_range(all entities, pinned host) + astre_gpu_set_views + astre_gpu_frame(ALL | GATHER) + "
                       "astre_gpu_read_counts + astre_gpu_read_draw_list (sorted keys + world matrices of the visible entities)"}
        # the case the mirror was built for: a dirty subset pushed as a slot list (astre_gpu_update_trs)
        rng = np.random.default_rng(7)
        for pct in (1, 10):
            m = max(n * pct // 100, 1)
            slots = np.sort(rng.choice(n, size=m, replace=False)).astype(np.uint32)
            sl_h = torch.from_numpy(slots).pin_memory().numpy()
            p_h, r_h, s_h = (torch.from_numpy(np.ascontiguousarray(a[slots])).pin_memory().numpy() for a in (s.pos, s.rot, s.scale))

            def dirty():
                ctx.update_trs(sl_h, p_h, r_h, s_h)
                set_views()
                ctx.frame(self.e2e_flags)
                return read_results()

            ms_d, nk_d = timed(dirty, steps)
            out[f"e2e_dirty_{pct}pct"] = {"value": self.total_entities / (ms_d * 1e-3), "unit": "entities/s", "ms_per_step": ms_d,
                                          "h2d_bytes_per_step": int(m * 44), "d2h_bytes_per_step": int(self.n_counts * 4 + nk_d * 72)}
        return out

    def parity_and_cpu(self, reps, want_cpu):
        """Replays this rank's shard with the CPU oracle: parity of counts / keys / world matrices, and (rank 0) the timing."""
        ol = _oracle()
        threads = max(host_threads() // self.world, 1)
        self.ctx.frame(self.gpu.FRAME_ALL)
        g_counts, g_keys = self.ctx.read_counts(), self.ctx.read_all_keys()
        g_world = self.ctx.read_world(0, self.n)
        dts = []
        if self.n <= 2_000_000:           # small frames: thread start-up would dominate a single cold call
            reps = max(reps, 10)
            for _ in range(2):
                cpu_port_frame(self.scene, threads)
        for _ in range(reps):
            dt, o_world, o_keys, o_counts, used = cpu_port_frame(self.scene, threads)
            dts.append(dt)
        ok_counts = bool(np.array_equal(g_counts, o_counts))
        ok_keys = bool(len(g_keys) == len(o_keys) and np.array_equal(g_keys, o_keys))
        ok_world = bool(np.array_equal(ol.bits(g_world), ol.bits(o_world)))
        offsets_ok = True
        if self.exchange:      # the exchanged table and the device-side offsets of the last step against the host formula
            self.step()
            table, off, tot = self.exchange.last()
            h_off, h_tot = self.gpu.global_offsets(table, self.rank, self.rank_major)
            offsets_ok = bool(np.array_equal(table[self.rank], o_counts) and np.array_equal(off, h_off) and np.array_equal(tot, h_tot))
        ok = ok_counts and ok_keys and ok_world and offsets_ok
        ranks_ok = int(self.max_over_ranks(1.0 if ok else 0.0, "sum"))
        parity = {"ranks_ok": ranks_ok, "ranks": self.world, "rank0": {"counts": ok_counts, "keys": ok_keys, "world_matrices": ok_world,
                                                                        "exchange_offsets": offsets_ok},
                  "keys_compared_rank0": int(len(o_keys)), "matrices_compared_rank0": int(self.n),
                  "checked": "per-view counts, the full sorted key list and every world matrix of each rank's shard, bit for bit "
                             "against oracle/astre_oracle.c (orc_frame)" + ("; all-gathered count table and device-side segment offsets "
                                                                            "against astre_global_offsets" if self.exchange else "")}
        cpu = None
        if want_cpu and self.rank == 0:
            cpu = {"value": self.n / float(np.mean(dts)), "unit": "entities/s", "cores": used, "kind": "port",
                   "sample": f"{reps} full frame(s) of rank 0's {self.n}-entity workload (oracle/astre_oracle.c, OpenMP, {used} threads)"}
        return parity, cpu


def run_workload(name, rank, local_rank, world, args, torch, dist, steps, warmup, e2e_steps, cpu_reps, headline):
    b = Bench(name, rank, local_rank, world, torch, dist, use_graph=not args.no_graph, exchange=args.exchange)
    try:
        ms_step, launches = b.device_resident(steps, warmup)
        kernel_ms = b.kernel_timing(min(steps, 20))
        out = {"workload": b.workload, "scaling": b.scaling, "value": b.total_entities / (ms_step * 1e-3), "unit": "entities/s",
               "ms_per_step": ms_step, "entities_per_gpu": b.n, "views": b.F, "visible_keys_rank0": int(b.total_keys),
               "sort_path": {0: "none", 1: "bucket", 2: "lsd-fallback"}[b.sort_path], "gpu_launches": launches,
               "frame_as_cuda_graph": bool(b.frame_flags & b.gpu.FRAME_GRAPH), "scaling_detail": b.scaling_detail,
               "roofline": b.roofline(kernel_ms, ms_step), "e2e": b.end_to_end(e2e_steps)}
        parity, cpu = b.parity_and_cpu(cpu_reps, not args.no_cpu_baseline)
        out["parity"], out["cpu_baseline"] = parity, cpu
        if headline:
            out["_bench"] = b
            b = None
        return out
    finally:
        if b is not None:
            b.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="astre_b200")
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5], help="make this BASELINE.json configuration the headline workload")
    ap.add_argument("--no-extra", action="store_true", help="skip config.extra_workloads")
    ap.add_argument("--entities", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: count boards (stores into peer memory, default) or an NCCL all-gather")
    ap.add_argument("--no-graph", action="store_true", help="launch the frame kernel by kernel instead of replaying one CUDA graph")
    args = ap.parse_args()
    if args.entities:
        globals()["N_PER_GPU"] = args.entities
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: astre_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    head_name = f"config{args.config}" if args.config else ("config3" if world == 1 else "config4w")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()     # clocks are sampled from warm-up to the end of the kernel-timing loop
    head = run_workload(head_name, rank, local_rank, world, args, torch, dist, args.steps, args.warmup,
                        max(3, min(args.steps, 10)), 3 if head_name in ("config3", "config4w", "config1", "config2") else 1, True)
    b = head.pop("_bench")
    # keep the GPU under the same load for a moment so that the 100 ms sampler sees it
    t_end = time.perf_counter() + 1.0
    while time.perf_counter() < t_end:
        for _ in range(50):
            b.ctx.frame(b.gpu.FRAME_ALL)
        b.ctx.sync()
    clocks = sampler.stop() if rank == 0 else None
    b.close()

    extras = {}
    if not args.no_extra and not args.config:
        names = (["config1"] if world == 1 else []) + ["config2", "config4", "config5"]
        for name in names:
            try:
                extras[name] = run_workload(name, rank, local_rank, world, args, torch, dist, min(args.steps, 20), 3, 3, 1, False)
            except Exception as e:      # an extra must never take the headline line down with it
                extras[name] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        n = head["entities_per_gpu"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "entities/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": head["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(head["workload"], n, head["views"], head["visible_keys_rank0"], world),
            "sort_path": head["sort_path"], "frame_as_cuda_graph": head["frame_as_cuda_graph"], "scaling_detail": head["scaling_detail"],
            "extra_workloads": extras,
            "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"], "parity": head["parity"],
            "gpu_launches": head["gpu_launches"], "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
