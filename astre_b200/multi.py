"""Multi-GPU orchestration: one process per GPU, entities sharded by contiguous range (flat scenes) or
by whole worlds (batched worlds); the only exchange on the path is an all-gather of per-shard visible
counts, from which every rank derives the offset of its segment in each global draw list
(SURVEY.md 8e).  Keys never leave their GPU.  Works over NCCL (device tensors) and gloo (CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

from . import gpu as _gpu


def shard_range(n, rank, world_size, align=4):
    """Contiguous entity range [lo, hi) of `rank`; boundaries are multiples of `align`."""
    per = -(-n // world_size)
    per = (per + align - 1) // align * align
    return min(rank * per, n), min((rank + 1) * per, n)


def shard_worlds(n_worlds, rank, world_size):
    """Whole worlds [lo, hi) of `rank` (batched independent worlds)."""
    per = -(-n_worlds // world_size)
    return min(rank * per, n_worlds), min((rank + 1) * per, n_worlds)


def shard_subtrees(parent, world_size):
    """Hierarchies shard by ROOT SUBTREE, so parent-before-child never crosses a GPU.
    Subtrees, in root order, are dealt to ranks in contiguous groups balanced by node count (to within
    one subtree).  Returns, per rank, (global entity indices ascending, parent as local slot)."""
    parent = np.asarray(parent, np.uint32)
    n = len(parent)
    idx = np.arange(n, dtype=np.int64)
    is_root = parent == _gpu.NO_PARENT
    anc = np.where(is_root, idx, parent.astype(np.int64))
    if n and (anc.min() < 0 or anc.max() >= n):
        raise ValueError("parent index out of range")
    for _ in range(64):                                   # pointer jumping: anc -> root in log2(depth) rounds
        nxt = anc[anc]
        if np.array_equal(nxt, anc):
            break
        anc = nxt
    else:
        raise ValueError("cyclic parent links")
    if n and not is_root[anc].all():
        raise ValueError("cyclic parent links")
    size = np.bincount(anc, minlength=n)                  # nodes per root (0 for non-roots)
    before = np.cumsum(size) - size                       # nodes of the subtrees that come earlier
    rank_of_root = np.minimum(before * world_size // max(n, 1), world_size - 1)
    rank_of = rank_of_root[anc]
    out = []
    for r in range(world_size):
        mine = np.flatnonzero(rank_of == r)
        local = np.full(n, -1, np.int64)
        local[mine] = np.arange(len(mine))
        p = parent[mine]
        lp = np.where(p == _gpu.NO_PARENT, np.int64(_gpu.NO_PARENT), local[np.where(p == _gpu.NO_PARENT, 0, p)])
        out.append((mine, lp.astype(np.uint32)))
    return out


def exchange_counts(counts, group=None):
    """All-gather of this shard's per-view visible counts.  `counts`: 1-D int32 tensor (device tensor
    under NCCL).  Returns a [world_size, n_views] tensor on the same device."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    if ws == 1:
        return counts.reshape(1, -1)
    out = torch.empty(ws * counts.numel(), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out, counts.contiguous(), group=group)
    return out.reshape(ws, -1)


def segment_offsets(all_counts, rank, rank_major=False):
    """Offset of `rank`'s segment of every view in the global draw list -- ordered (view, rank, key) for flat scenes
    sharded by entity range, rank-major (= world-major) for batched worlds sharded by whole worlds -- and the totals."""
    a = all_counts.detach().cpu().numpy().astype(np.uint32)
    return _gpu.global_offsets(a, rank, rank_major)


def device_view(ptr, n, typestr, device):
    """Zero-copy torch view of `n` elements of library-owned device memory."""
    class _View:
        __cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_View(), device=device)


class CountExchange:
    """The per-frame exchange of SURVEY.md 8e, kept OFF the frame's critical path.

    The library mirrors every frame's per-view counts into one of two device slots (astre_gpu_set_counts_mirror; the
    frame's last kernel writes them, so the frame stays one CUDA graph).  After a frame the caller's stream only
    records an event; everything else runs on a side stream while the next frames compute: NCCL all-gather of the
    slot -> [world_size, n_views] table -> astre_gpu_global_offsets_dev (this rank's segment offsets and the per-view
    totals, on the device).  A slot is rewritten two frames later, and only after its exchange has finished (an
    event wait that is already satisfied in steady state).  `finish()` joins the side stream back (end of a timed
    region, or before reading).  No host round trip anywhere."""

    def __init__(self, ctx, n_counts, group=None, device=None, rank_major=False):
        self.ctx, self.group, self.n, self.rank_major = ctx, group, int(n_counts), bool(rank_major)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.slot = [torch.zeros(self.n, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.table = [torch.zeros((self.world, self.n), dtype=torch.int32, device=self.device) for _ in range(2)]
        self.offsets = [torch.zeros(self.n, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.totals = [torch.zeros(self.n, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.side = torch.cuda.Stream(device=self.device)
        self.framed = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.busy = [False, False]
        self.last_slot = None
        ctx.set_counts_mirror(self.slot[0].data_ptr(), self.slot[1].data_ptr())

    def close(self):
        self.finish()
        torch.cuda.synchronize(self.device)
        self.ctx.set_counts_mirror(0, 0)


    def before_frame(self):
        """Call before ctx.frame(): the slot the coming frame writes must not be in use by an exchange any more."""
        p = (self.ctx.frame_index() + 1) & 1
        if self.busy[p]:
            torch.cuda.current_stream(self.device).wait_event(self.done[p])
            self.busy[p] = False

    def after_frame(self):
        """Call after ctx.frame() on the frame's stream: hands the frame's counts to the side stream."""
        p = self.ctx.frame_index() & 1
        main = torch.cuda.current_stream(self.device)
        self.framed[p].record(main)
        self.side.wait_event(self.framed[p])
        with torch.cuda.stream(self.side):
            self._exchange(p)
            self.done[p].record(self.side)
        self.busy[p] = True
        self.last_slot = p
        return p

    def _exchange(self, p):
        """On the side stream: slot p -> table p, offsets p, totals p."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.table[p].view(-1), self.slot[p], group=self.group)
        else:
            self.table[p].view(-1).copy_(self.slot[p], non_blocking=True)
        self.ctx.global_offsets_dev(self.table[p].data_ptr(), self.world, self.rank, self.n,
                                    self.offsets[p].data_ptr(), self.totals[p].data_ptr(), self.side.cuda_stream,
                                    rank_major=self.rank_major)

    def finish(self):
        torch.cuda.current_stream(self.device).wait_stream(self.side)

    def last(self):
        """(all_counts [world, n], offsets [n], totals [n]) of the most recent frame, on the host (synchronises)."""
        p = self.last_slot
        self.done[p].synchronize()
        return (self.table[p].cpu().numpy().astype(np.uint32), self.offsets[p].cpu().numpy().astype(np.uint64),
                self.totals[p].cpu().numpy().astype(np.uint64))


class PeerCountExchange(CountExchange):
    """The same exchange with no collective library on the path: count boards (astre_gpu_count_board_*).  Every rank's
    board lives in its own HBM and is mapped by the peers through CUDA IPC; per frame ONE one-block kernel on the side
    stream stores this rank's counts into all boards over NVLink, waits for everybody's, and writes the table, this
    rank's segment offsets and the per-view totals -- all-gather and offsets fused over peer memory.  The frame itself
    only mirrors its counts into a slot, exactly as for CountExchange.  `boards` / `rank` / `world`: device pointers of
    all ranks' boards when the ranks live in one process (tests); otherwise the IPC handles are exchanged once."""

    def __init__(self, ctx, n_counts, group=None, device=None, rank_major=False, boards=None, rank=None, world=None):
        super().__init__(ctx, n_counts, group=group, device=device, rank_major=rank_major)
        if rank is not None:
            self.rank, self.world = int(rank), int(world)
            self.table = [torch.zeros((self.world, self.n), dtype=torch.int32, device=self.device) for _ in range(2)]
        self.imports = []
        if boards is None:          # one process per GPU: exchange the IPC handles once
            handle, own = ctx.count_board_create(self.n)
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, (self.n, handle), group=group)
                if any(h[0] != self.n for h in handles):
                    raise ValueError(f"count boards need the same number of counts on every rank, got {[h[0] for h in handles]}")
            boards = [own if r == self.rank else ctx.import_handle(handles[r][1]) for r in range(self.world)]
            self.imports = [b for r, b in enumerate(boards) if r != self.rank]
        ctx.count_board_attach(self.world, self.rank, boards)

    def close(self):
        super().close()
        exchanges, error = self.ctx.count_board_status()
        if error:
            raise RuntimeError(f"count board: rank {error - 1} did not answer in time (exchange {exchanges + 1})")
        if self.imports and dist.is_initialized():
            dist.barrier(group=self.group)         # nobody unmaps a board a peer may still be storing into
        for p in self.imports:
            self.ctx.release_import(p)
        self.imports = []

    def _exchange(self, p):
        self.ctx.count_board_exchange(self.slot[p].data_ptr(), self.n, self.offsets[p].data_ptr(), self.totals[p].data_ptr(),
                                      self.table[p].data_ptr(), stream=self.side.cuda_stream, rank_major=self.rank_major)


class GlobalDrawList:
    """f3: one globally key-sorted draw list out of the ranks' sorted lists, ordered
    (key >> 26, rank, local slot) = class, depth, global entity order (flat scenes sharded by range).

    mode "p2p":  every rank publishes its list at a fixed device address, the addresses are exchanged
                 once as CUDA IPC handles, and the merge kernel reads the peers' lists over NVLink
                 itself -- the exchange is fused into the merge, no gather buffer, and with
                 `sliced=True` a rank pulls only the keys of its own slice of the global list.
    mode "nccl": all-gather of the (padded) lists, then the same merge kernel on the gathered rows
                 (the library-collective baseline the p2p mode is measured against).
    Per frame the only host round trip is reading the all-gathered per-view counts."""

    def __init__(self, ctx, n_views, max_keys, mode="p2p", group=None, device=None):
        assert mode in ("p2p", "nccl")
        self.ctx, self.mode, self.group, self.n_views, self.max_keys = ctx, mode, group, n_views, int(max_keys)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.counts_dev = device_view(ctx.device_counts_ptr(), n_views, "<i4", self.device)
        self.local = [ctx.device_published_keys_ptr(p) for p in (0, 1)]
        self.local_view = [device_view(self.local[p], self.max_keys, "<i8", self.device) for p in (0, 1)]
        self.peer = [[None, None] for _ in range(self.world)]
        self.peer[self.rank] = list(self.local)
        self._gathered = None
        self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._epoch = 0
        self._bind()
        if mode == "p2p" and self.world > 1:
            mine = [self.max_keys, ctx.export_handle(0), ctx.export_handle(1)]
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            if any(h[0] != self.max_keys for h in handles):
                # the publish buffer layout (keys | samples | cut table) is derived from max_keys on every rank
                raise ValueError(f"peer-to-peer merge needs the same max_keys on every rank, got {[h[0] for h in handles]}")
            for r in range(self.world):
                if r != self.rank:
                    self.peer[r] = [ctx.import_handle(handles[r][1]), ctx.import_handle(handles[r][2])]

    def _bind(self):
        """The collectives below are issued by torch on ITS current stream; the flow is only correct if the library's
        frame, publish and merge kernels are ordered on that same stream.  Bind the context to it (a change of
        stream synchronises the old one, so work already enqueued stays ordered)."""
        cur = torch.cuda.current_stream(self.device).cuda_stream
        if self.ctx.bound_stream != cur:
            self.ctx.set_stream(cur)
        return cur

    def build_enqueue(self, frame_index, sliced=False, out_cap=0, stream=None):
        """Peer-to-peer flow with the key counts left on the device: nothing here waits for the GPU, so frames
        and merges can be enqueued back to back.  The list is in the library's merge buffers afterwards
        (ctx.merged_count() / ctx.read_merged synchronise)."""
        assert self.mode == "p2p" or self.world == 1
        cur = self._bind()
        assert stream in (None, cur), "pass the torch current stream (or None): the NCCL calls run on it"
        parity = frame_index & 1
        self.ctx.publish_keys(parity, stream)
        self._all_counts = exchange_counts(self.counts_dev, self.group)          # [G, F] int32, stays on the device
        ptrs = [self.peer[r][parity] for r in range(self.world)]
        self.ctx.merge_prepare_dev(self.rank, ptrs, self._all_counts.data_ptr(), self.n_views, stream)
        if self.world > 1:
            dist.all_reduce(self._token, group=self.group)                       # device-side barrier
        self.ctx.merge_published_dev(ptrs, self.rank if sliced else 0, self.world if sliced else 0, out_cap, stream)

    def build_flags(self, sliced=False, out_cap=0, stream=None):
        """Peer-to-peer flow without NCCL: counts and ready flags travel in the publish buffers and are polled
        over NVLink by the library's own kernels.  Call once per frame on every rank (after ctx.frame());
        the epoch counter lives here and must advance in step on all ranks."""
        assert self.mode == "p2p" or self.world == 1
        self._bind()
        self._epoch += 1
        p0 = [self.peer[r][0] for r in range(self.world)]
        p1 = [self.peer[r][1] for r in range(self.world)]
        self.ctx.merge_p2p(self.rank, p0, p1, self._epoch, self.rank if sliced else 0, self.world if sliced else 0, out_cap, stream)

    def build(self, frame_index, sliced=False, stream=None):
        """Call after ctx.frame().  Returns (all_counts [G, F] on the host, out_first, out_count): the
        merged keys / sources of positions [out_first, out_first + out_count) are in the library's
        merge buffers (ctx.read_merged / ctx.device_merged_ptrs)."""
        cur = self._bind()
        assert stream in (None, cur), "pass the torch current stream (or None): the NCCL calls run on it"
        parity = frame_index & 1
        self.ctx.publish_keys(parity, stream)
        # the counts all-gather is stream-ordered after the publish copy on every rank: once it has
        # completed here, every peer's published list is complete
        all_counts = exchange_counts(self.counts_dev, self.group).cpu().numpy().astype(np.uint64)
        totals = all_counts.sum(axis=1)
        total = int(totals.sum())
        if sliced:
            per = -(-total // self.world)
            first = min(self.rank * per, total)
            count = min(per, total - first)
        else:
            first, count = 0, total
        if self.mode == "p2p" or self.world == 1:
            ptrs = [self.peer[r][parity] for r in range(self.world)]
            # each rank searches its own list only; the cut tables are read by the peers afterwards
            self.ctx.merge_prepare(self.rank, ptrs, totals, stream)
            if self.world > 1:
                dist.all_reduce(self._token, group=self.group)      # device-side barrier, no host sync
            self.ctx.merge_published(ptrs, totals, first, count, stream)
            return all_counts, first, count
        else:
            cap = max(int(totals.max()), 1)
            if self._gathered is None or self._gathered.numel() < self.world * cap:
                self._gathered = torch.empty(self.world * (cap + cap // 4), dtype=torch.int64, device=self.device)
            rows = self._gathered[: self.world * cap]
            dist.all_gather_into_tensor(rows, self.local_view[parity][:cap], group=self.group)
            ptrs = [rows.data_ptr() + 8 * cap * r for r in range(self.world)]
        self.ctx.merge_lists(ptrs, totals, first, count, stream)
        return all_counts, first, count
