"""Multi-GPU sharding of the visibility pipeline. One process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests). The reference is single-GPU; this is new.

Two ways the path shards:
  1. by VIEW (shadow cascades, many-camera batches): independent units, no data-path collective —
     `views_for_rank`.
  2. by MESHLET RANGE of one huge view (C3): the entity-draw array is cut into `world` contiguous ranges with
     equal LOD-0 meshlet sums (boundaries multiples of 32 draws so visibility words stay disjoint); each rank
     culls its range; the exchange step is (a) broadcast of the small Hi-Z pyramid from the rank that owns the
     depth buffer and (b) an all-gather of survivor counts followed by an all-gather of the survivor lists into
     the rank-major draw list — `partition_draws`, `broadcast_pyramid`, `exchange_survivors`.
"""
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

DRAW_BYTES = 28


def views_for_rank(n_views: int, rank: int, world: int) -> List[int]:
    """View v is culled by rank v mod world."""
    return [v for v in range(n_views) if v % world == rank]


def partition_draws(meshlet_counts: Sequence[int], world: int, weights: Sequence[float] = None) -> List[Tuple[int, int]]:
    """Cuts draws [0, N) into `world` contiguous ranges of (nearly) equal meshlet sums — or sums proportional to `weights` —
    whose interior boundaries are multiples of 32 draws (entity visibility words are indexed by draw/32). Ranges may be empty."""
    c = np.asarray(meshlet_counts, dtype=np.int64)
    n = len(c)
    prefix = np.concatenate([[0], np.cumsum(c)])
    total = int(prefix[-1])
    w = np.ones(world) if weights is None else np.asarray(weights, dtype=np.float64)
    assert len(w) == world and (w >= 0).all() and w.sum() > 0
    share = np.concatenate([[0.0], np.cumsum(w) / w.sum()])
    cuts = [0]
    for k in range(1, world):
        target = total * share[k]
        i = int(np.searchsorted(prefix, target, side="left"))
        i = int(round(i / 32.0)) * 32
        i = min(max(i, cuts[-1]), n)
        cuts.append(i)
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def broadcast_pyramid(texels: torch.Tensor, src: int = 0, group=None) -> None:
    """The pyramid is one contiguous f32 allocation (5.6 MB at 1080p, 22.4 MB at 4K): one broadcast."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(texels, src=src, group=group)


def exchange_survivors(local_draw_buffer: torch.Tensor, out_buffer: torch.Tensor = None, group=None):
    """local_draw_buffer: uint8 MeshletDrawCommandBuffer of this rank (u32 count @0, 28-byte commands @4).
    Returns (global MeshletDrawCommandBuffer in rank-major order, per-rank counts). Every rank gets the full list."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = local_draw_buffer.device
    count = local_draw_buffer[:4].view(torch.int32).clone()
    if world == 1:
        n = int(count.item())
        return local_draw_buffer[:4 + DRAW_BYTES * n], [n]
    counts = torch.zeros(world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(counts, count, group=group)
    cap = (local_draw_buffer.numel() - 4) // DRAW_BYTES
    counts_h = [min(int(v), cap) for v in counts.cpu().tolist()]      # an overflowed rank holds only `capacity` commands (equal capacities)
    total, mx = sum(counts_h), max(counts_h)
    rank = dist.get_rank(group)
    # NCCL has no all-gather-v: pad every rank's list to the longest one
    send = torch.zeros(max(mx, 1) * DRAW_BYTES, dtype=torch.uint8, device=dev)
    send[:counts_h[rank] * DRAW_BYTES] = local_draw_buffer[4:4 + counts_h[rank] * DRAW_BYTES]
    gathered = torch.empty(world * send.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, send, group=group)
    if out_buffer is None or out_buffer.numel() < 4 + DRAW_BYTES * total:
        out_buffer = torch.empty(4 + DRAW_BYTES * total, dtype=torch.uint8, device=dev)
    out_buffer[:4] = torch.tensor([total], dtype=torch.int32, device=dev).view(torch.uint8)
    off = 4
    for r in range(world):
        nb = counts_h[r] * DRAW_BYTES
        out_buffer[off:off + nb] = gathered[r * send.numel():r * send.numel() + nb]
        off += nb
    return out_buffer[:4 + DRAW_BYTES * total], counts_h


class PeerExchange:
    """Survivor exchange over NVLink peer stores instead of a padded NCCL all-gather.

    Every rank owns one output MeshletDrawCommandBuffer in cudaMalloc'd memory shared through CUDA IPC
    (orbit_peer_alloc / orbit_peer_open). After the per-rank counts are all-gathered (4 bytes each, the only
    collective), each rank's `orbit_draws_scatter` kernel stores its commands straight into EVERY rank's output
    buffer at command index sum(counts[:rank]) — remote stores travel over NVLink — so every rank ends up with the
    full rank-major list without staging or padding. A barrier closes the step."""

    def __init__(self, context, capacity_draws, group=None):
        import ctypes as C
        from . import _lib
        self.C, self.lib, self.context, self.group = C, _lib.lib(), context, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.capacity = int(capacity_draws)
        self.bytes = 4 + DRAW_BYTES * self.capacity
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(self.lib.orbit_peer_alloc(context._h, self.bytes, C.byref(ptr), handle), "orbit_peer_alloc")
        self.local_ptr = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.peer_ptrs = []
        for r in range(self.world):
            if r == self.rank:
                self.peer_ptrs.append(self.local_ptr)
            else:
                q = C.c_void_p()
                _lib.check(self.lib.orbit_peer_open(context._h, handles[r], C.byref(q)), "orbit_peer_open")
                self.peer_ptrs.append(q.value)
        self.counts = torch.zeros(self.world, dtype=torch.int32, device=context.device)
        dist.barrier(group=group)

    def exchange(self, local_draw_buffer, root=None):
        """Returns (total count, per-rank counts); the assembled list is in self.local_ptr (see read()).
        root=None: every rank receives the full list (all-gather). root=r: only rank r does (gather) — what a frame
        whose draws are submitted by one GPU needs; each rank then stores its commands once instead of world times."""
        C = self.C
        dist.all_gather_into_tensor(self.counts, local_draw_buffer[:4].view(torch.int32), group=self.group)
        counts = [min(int(v), self.capacity) for v in self.counts.cpu().tolist()]     # an overflowed rank stored only `capacity` commands
        total, first = sum(counts), sum(counts[:self.rank])
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for k in range(self.world):
            r = (self.rank + k) % self.world       # staggered destinations: no rank is everybody's first target
            if root is not None and r != root:
                continue
            rc = self.lib.orbit_draws_scatter(self.context._h, C.c_void_p(local_draw_buffer.data_ptr()), self.capacity, C.c_void_p(self.peer_ptrs[r]),
                                              first, total, self.capacity, stream)
            if rc:
                raise RuntimeError("orbit_draws_scatter: %d" % rc)
        dist.barrier(group=self.group)   # every rank's stores into my buffer are complete and visible
        return total, counts

    def fence(self):
        """Closing fence of exchange_async(..., fence=False): completes on a rank only after every rank's stores were issued and flushed."""
        if not hasattr(self, "_fence"):
            self._fence = torch.zeros(1, dtype=torch.int32, device=self.context.device)
        dist.all_reduce(self._fence, group=self.group)

    def exchange_async(self, local_draw_buffer, root=None, fence=True):
        """Same exchange without any host round trip: the all-gathered counts stay on the device and
        `orbit_draws_scatter_ranked` derives each rank's offset from them; a 4-byte all-reduce enqueued behind the
        stores is the closing barrier. Everything is stream-ordered, so a sharded frame can be enqueued back to back.
        Returns the device tensor of per-rank counts (read it after synchronising to learn the total)."""
        C = self.C
        dist.all_gather_into_tensor(self.counts, local_draw_buffer[:4].view(torch.int32), group=self.group)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for k in range(self.world):
            r = (self.rank + k) % self.world
            if root is not None and r != root:
                continue
            rc = self.lib.orbit_draws_scatter_ranked(self.context._h, C.c_void_p(local_draw_buffer.data_ptr()), self.capacity, C.c_void_p(self.peer_ptrs[r]),
                                                     C.c_void_p(self.counts.data_ptr()), self.rank, self.world, self.capacity, stream)
            if rc:
                raise RuntimeError("orbit_draws_scatter_ranked: %d" % rc)
        if fence:
            self.fence()
        return self.counts

    def read(self, total):
        """Copies the assembled MeshletDrawCommandBuffer (count + total commands) into a torch tensor."""
        out = torch.empty(4 + DRAW_BYTES * total, dtype=torch.uint8, device=self.context.device)
        self.lib.orbit_device_copy(self.C.c_void_p(out.data_ptr()), self.C.c_void_p(self.local_ptr), out.numel(),
                                   self.C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out

    def clear(self):
        """Zeroes this rank's assembled list (tests: proves that a following exchange really delivers it)."""
        z = torch.zeros(self.bytes, dtype=torch.uint8, device=self.context.device)
        self.lib.orbit_device_copy(self.C.c_void_p(self.local_ptr), self.C.c_void_p(z.data_ptr()), z.numel(),
                                   self.C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()

    def close(self):
        for r, q in enumerate(self.peer_ptrs):
            if r != self.rank:
                self.lib.orbit_peer_close(self.context._h, self.C.c_void_p(q))
        self.lib.orbit_peer_free(self.context._h, self.C.c_void_p(self.local_ptr))


class MaskExchange:
    """Survivor exchange of a sharded view in its compact form: every rank ships one 16-byte entry per dispatch record
    ({draw mask, entity, meshlet offset, 1}, written by orbit_meshlet_test) to the ROOT rank — the GPU that submits the draws —
    by NVLink peer stores into ITS OWN region of the root's entry array, plus its record count into its slot of the root's
    count words (orbit_record_masks_put): no rank needs another rank's count, so the only collective is the closing 4-byte
    fence. The root reads the regions in rank order (rank-major = canonical record order), recounts and emits the 28-byte
    commands itself (orbit_draws_from_masks). C3: 28 MB cross the switch instead of 229 MB, and the emission (a local
    HBM-bound kernel) no longer waits for a single GPU's NVLink ingress."""

    HEADER_BYTES = 256      # the root's count words live in front of the regions

    def __init__(self, context, capacity_records_rank, root=0, group=None):
        import ctypes as C
        from . import _lib
        self.C, self.lib, self.context, self.group, self.root = C, _lib.lib(), context, group, root
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert self.world <= 16
        cap = torch.tensor([int(capacity_records_rank)], dtype=torch.int64, device=context.device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)          # one region size for every rank
        self.rank_capacity = max(int(cap.item()), 1)
        self.local_masks = torch.zeros(16 * self.rank_capacity, dtype=torch.uint8, device=context.device)
        nbytes = self.HEADER_BYTES + 16 * self.rank_capacity * self.world if self.rank == root else 256
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(self.lib.orbit_peer_alloc(context._h, nbytes, C.byref(ptr), handle), "orbit_peer_alloc")
        self.local_ptr = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        if self.rank == root:
            self.root_ptr = self.local_ptr
        else:
            q = C.c_void_p()
            _lib.check(self.lib.orbit_peer_open(context._h, handles[root], C.byref(q)), "orbit_peer_open")
            self.root_ptr = q.value
        self._fence = torch.zeros(1, dtype=torch.int32, device=context.device)
        dist.barrier(group=group)

    def exchange(self, dispatch_buffer):
        """Enqueued on the current stream: this rank's entries + count -> its region on the root, then the closing fence."""
        C = self.C
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        region = self.root_ptr + self.HEADER_BYTES + 16 * self.rank_capacity * self.rank
        rc = self.lib.orbit_record_masks_put(self.context._h, C.c_void_p(self.local_masks.data_ptr()), C.c_void_p(dispatch_buffer.data_ptr()),
                                             self.rank_capacity, C.c_void_p(region), C.c_void_p(self.root_ptr + 4 * self.rank), stream)
        if rc:
            raise RuntimeError("orbit_record_masks_put: %d" % rc)
        dist.all_reduce(self._fence, group=self.group)     # completes on the root only after every rank's stores were issued and flushed

    def expand(self, scene_buffers, draw_buffer, capacity_draws):
        """Root only, after exchange(): the ranks' regions -> MeshletDrawCommandBuffer."""
        C = self.C
        rc = self.lib.orbit_draws_from_masks(self.context._h, C.byref(scene_buffers), C.c_void_p(self.local_ptr + self.HEADER_BYTES),
                                             self.rank_capacity, C.c_void_p(self.local_ptr), self.world,
                                             C.c_void_p(draw_buffer.data_ptr()), int(capacity_draws),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            raise RuntimeError("orbit_draws_from_masks: %d" % rc)

    def close(self):
        if self.rank != self.root:
            self.lib.orbit_peer_close(self.context._h, self.C.c_void_p(self.root_ptr))
        self.lib.orbit_peer_free(self.context._h, self.C.c_void_p(self.local_ptr))


class PyramidBroadcast:
    """The depth pyramid from the rank that built it to every other rank WITHOUT a collective: the pyramids live in
    peer-mapped memory, the source rank scatters one chunk to each other rank (orbit_peer_put: NVLink stores + a completion
    flag on the receiver), every rank forwards the chunk it received to the remaining ranks, and waits for their chunks'
    flags (orbit_peer_wait). Each GPU sends and receives ~one pyramid's worth of bytes through the switch instead of the
    source sending world-1 copies or a ring passing it hand to hand: the NCCL broadcast of the 22.4 MB 4K pyramid took
    150-340 us on 8 GPUs (profiles/r2_c3_timeline_n8.txt) and was the largest item of the sharded frame.
    The caller guarantees what a collective would: no rank calls broadcast() for frame N+1 before every rank has finished
    reading the pyramid of frame N (ShardedView.step_best: its closing fence)."""

    FLAG_BYTES = 256     # word 0: "my chunk has arrived" (written by the source), word 1+q: "rank q's chunk has arrived"

    def __init__(self, context, pyramid, src=0, group=None):
        import ctypes as C
        from . import _lib, layouts as L
        self.C, self.L, self.lib, self.context, self.group, self.src = C, L, _lib.lib(), context, group, src
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert 2 <= self.world <= 16
        self.pyr_bytes = (4 * int(pyramid.info.total_texels) + 15) // 16 * 16
        nbytes = self.pyr_bytes + self.FLAG_BYTES
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(self.lib.orbit_peer_alloc(context._h, nbytes, C.byref(ptr), handle), "orbit_peer_alloc")
        self.local = ptr.value
        pyramid.rebind_external(self.local)
        pyramid.texels.zero_()
        class _Mem:
            def __init__(self, p, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(p), False), "version": 2}
        self._flag_mem = _Mem(self.local + self.pyr_bytes, self.FLAG_BYTES // 4)
        self.flags = torch.as_tensor(self._flag_mem, device=context.device)
        self.flags.zero_()
        torch.cuda.synchronize(context.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.peers = []
        for r in range(self.world):
            if r == self.rank:
                self.peers.append(self.local)
            else:
                q = C.c_void_p()
                _lib.check(self.lib.orbit_peer_open(context._h, handles[r], C.byref(q)), "orbit_peer_open")
                self.peers.append(q.value)
        # chunk k (16-byte aligned) belongs to the k-th rank that is not the source
        others = [r for r in range(self.world) if r != src]
        per = (self.pyr_bytes // 16 + len(others) - 1) // len(others) * 16
        self.chunk = {r: (min(i * per, self.pyr_bytes), min((i + 1) * per, self.pyr_bytes)) for i, r in enumerate(others)}
        self.epoch = 0
        dist.barrier(group=group)

    def _put(self, transfers):
        C = self.C
        arr = (self.L.PeerPut * len(transfers))()
        for a, (src, dst, nbytes, flag) in zip(arr, transfers):
            a.src, a.dst, a.bytes, a.dst_flag = src, dst, nbytes, flag
        rc = self.lib.orbit_peer_put(self.context._h, arr, len(transfers), self.epoch, C.c_void_p(torch.cuda.current_stream(self.context.device).cuda_stream))
        if rc:
            raise RuntimeError("orbit_peer_put: %d" % rc)

    def _wait(self, first_word, n):
        C = self.C
        rc = self.lib.orbit_peer_wait(self.context._h, C.c_void_p(self.local + self.pyr_bytes + 4 * first_word), n, 1, self.epoch,
                                      C.c_void_p(torch.cuda.current_stream(self.context.device).cuda_stream))
        if rc:
            raise RuntimeError("orbit_peer_wait: %d" % rc)

    def broadcast(self):
        """Enqueued on the current stream of every rank; on the source rank after the kernel that wrote the pyramid."""
        self.epoch += 1
        fo = self.pyr_bytes
        if self.rank == self.src:
            self._put([(self.local + b, self.peers[r] + b, e - b, self.peers[r] + fo) for r, (b, e) in self.chunk.items() if e > b])
            return
        b, e = self.chunk[self.rank]
        # my own slot and the source's are satisfied by definition, so the final wait can take all `world` slots at once
        self.flags[1 + self.rank] = self.epoch
        self.flags[1 + self.src] = self.epoch
        if e > b:
            self._wait(0, 1)                                       # my chunk has arrived from the source
        fwd = [(self.local + b, self.peers[q] + b, e - b, self.peers[q] + fo + 4 * (1 + self.rank))
               for q in range(self.world) if q not in (self.src, self.rank)]
        if fwd:
            self._put(fwd)                                         # (an empty chunk still delivers its flag)
        self._wait(1, self.world)                                  # every other rank's chunk has arrived

    def close(self):
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.lib.orbit_peer_close(self.context._h, self.C.c_void_p(p))
        self.lib.orbit_peer_free(self.context._h, self.C.c_void_p(self.local))


class ShardedView:
    """One huge view culled by `world` GPUs (BASELINE config C3). Rank 0 owns the depth buffer and builds the
    pyramid; every rank culls its entity-draw range with the CUDA passes; survivors are all-gathered."""

    def __init__(self, context, scene, view, depth_np, rank, world, root_weight=None):
        """`root_weight`: rank 0's share of the meshlets relative to the other ranks' 1.0. Rank 0 also builds the pyramid and
        emits both command lists (the early list's 229 MB of stores at C3 overlap its late test) — a fixed ~250 us of extra work,
        so an equal share makes it the straggler once the per-rank test time is of that order (4 GPUs, equal shares: late test
        311 us against 200 us on the others, profiles/r2_c3_timeline_n4.txt; 8 GPUs, frame end by the timeline tool: 461 us at
        0.75, 445 at 0.5, 444 at 0.3), while on 2 GPUs half a share only moves the work to rank 1 (C3 with exchange: 769 us
        with equal shares, 792 at 0.5). Default: max(0.25, 1.03 - 0.066 * world) — 0.9 / 0.77 / 0.5 on 2 / 4 / 8 GPUs;
        ORBIT_ROOT_WEIGHT overrides."""
        import os
        from . import frame
        self.frame = frame
        self.context, self.view, self.rank, self.world = context, view, rank, world
        lod0 = scene.mesh_infos["mesh_lods"][:, 0, 1][scene.draws["mesh_index"]]
        self._lod0_records = (lod0.astype(np.int64) + 31) // 32
        if root_weight is None:
            root_weight = float(os.environ.get("ORBIT_ROOT_WEIGHT", "%.3f" % max(0.25, 1.03 - 0.066 * world)))
        self.ranges = partition_draws(lod0, world, [root_weight] + [1.0] * (world - 1) if world > 1 else None)
        b, e = self.ranges[rank]
        if b == e:            # empty range: keep the launch legal
            e = b
        self.dscene = frame.DeviceScene.upload(context, scene, draw_begin=b, draw_end=e if e > b else b)
        self.empty = (e <= b)
        self.vstate = frame.ViewState(context, self.dscene, (view.width, view.height))
        self.depth = torch.from_numpy(depth_np).to(context.device) if rank == 0 else torch.zeros(
            (view.height, view.width), dtype=torch.float32, device=context.device)
        self.prepared = frame.PreparedFrame(context, self.dscene, self.vstate, view, self.depth)

    def enable_peer_exchange(self, capacity_draws):
        """Use NVLink peer stores (PeerExchange) for the survivor exchange instead of the padded NCCL all-gather."""
        self.peer_early = PeerExchange(self.context, capacity_draws)
        self.peer_late = PeerExchange(self.context, capacity_draws)

    def step_overlapped(self, root=0):
        """The same frame with the EARLY list's exchange overlapped with the rest of the frame: the early survivors are final
        as soon as the early pass ends, so their counts all-gather and peer stores run on a side stream while this rank builds /
        receives the pyramid and runs the late pass; the late list follows, and ONE fence closes both. (The fence has to come
        last: torch keeps one NCCL stream per group, so a fence issued right after the early stores would hold the pyramid
        broadcast back until those stores are done.) root=None: every rank receives both lists."""
        pf = self.prepared
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream()
        if not self.empty:
            pf.entity(False); pf.meshlet(False)
        else:
            pf.early_draws[:4].zero_()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            c_e = self.peer_early.exchange_async(pf.early_draws, root, fence=False)
        if self.rank == 0:
            pf.hiz()
        broadcast_pyramid(self.vstate.depth_pyramid.texels, src=0)
        if not self.empty:
            pf.entity(True); pf.meshlet(True)
        else:
            pf.late_draws[:4].zero_()
        c_l = self.peer_late.exchange_async(pf.late_draws, root, fence=False)
        main.wait_stream(self._side)
        self.peer_late.fence()
        return c_e, c_l

    def step(self, exchange=True):
        """One two-pass frame: early cull (local range) -> Hi-Z on rank 0 + broadcast -> late cull -> survivor exchange:
        True = padded NCCL all-gather, "peer" = NVLink peer stores to every rank, "gather" = peer stores to rank 0 only,
        False = none."""
        pf = self.prepared
        if not self.empty:
            pf.entity(False); pf.meshlet(False)
        if self.rank == 0:
            pf.hiz()
        broadcast_pyramid(self.vstate.depth_pyramid.texels, src=0)
        if not self.empty:
            pf.entity(True); pf.meshlet(True)
        else:
            pf.early_draws[:4].zero_(); pf.late_draws[:4].zero_()
        if not exchange:
            return None
        if exchange in ("peer_async", "gather_async"):
            root = 0 if exchange == "gather_async" else None
            return self.peer_early.exchange_async(pf.early_draws, root), self.peer_late.exchange_async(pf.late_draws, root)
        if exchange in ("peer", "gather"):
            root = 0 if exchange == "gather" else None
            n_e, _ = self.peer_early.exchange(pf.early_draws, root)
            n_l, _ = self.peer_late.exchange(pf.late_draws, root)
            return n_e, n_l
        early, _ = exchange_survivors(pf.early_draws)
        late, _ = exchange_survivors(pf.late_draws)
        return early, late

    # ---- the sharded frame with the compact exchange (what bench.py times as "including the exchange") --------------------
    def enable_mask_exchange(self, capacity_records_total, capacity_draws_total):
        """Record-entry exchange (MaskExchange) for both lists; the early list's exchange gets its own process group (its own
        NCCL stream), so it overlaps the pyramid broadcast and the late pass instead of queueing behind them, and runs — with
        rank 0's emission of the early list — on a low-priority stream, so that it fills the gaps of the critical path (Hi-Z,
        broadcast, late pass on a high-priority stream) instead of delaying it."""
        self.side_group = dist.new_group()
        lod0 = self._lod0_records
        b, e = self.ranges[self.rank]
        rcap_rank = int(lod0[b:e].sum())                          # records this rank can produce (every entity at LOD 0)
        self.mx_early = MaskExchange(self.context, rcap_rank, root=0, group=self.side_group)
        self.mx_late = MaskExchange(self.context, rcap_rank, root=0)
        self.total_dcap = int(capacity_draws_total)
        if self.rank == 0:
            self.gathered_early = torch.zeros(4 + DRAW_BYTES * self.total_dcap, dtype=torch.uint8, device=self.context.device)
            self.gathered_late = torch.zeros(4 + DRAW_BYTES * self.total_dcap, dtype=torch.uint8, device=self.context.device)
        self._side = torch.cuda.Stream()                          # default = lowest priority
        self._crit = torch.cuda.Stream(priority=-1)
        # the pyramid travels by one-sided NVLink puts (PyramidBroadcast) unless ORBIT_NCCL_PYRAMID=1 asks for the NCCL broadcast
        import os
        self.pyr_bcast = None
        if self.world > 1 and not os.environ.get("ORBIT_NCCL_PYRAMID"):
            self.pyr_bcast = PyramidBroadcast(self.context, self.vstate.depth_pyramid, src=0)

    def best_exchange_name(self):
        return ("16-byte record entries {draw mask, entity, meshlet offset} to rank 0 by NVLink peer stores (device-side counts), "
                "rank 0 emits the commands of the combined list; the early list's exchange + emission overlap Hi-Z, its broadcast and the late pass; "
                + ("pyramid: scatter + forward by one-sided NVLink puts with completion flags" if self.pyr_bcast is not None else "pyramid: NCCL broadcast"))

    def step_best(self, marks=None):
        """early cull (test only) -> [side stream: entries -> rank 0, rank 0 emits the early list] || Hi-Z on rank 0 + broadcast ->
        late cull (test only) -> entries -> rank 0, rank 0 emits the late list. `marks`: a dict that receives CUDA events recorded
        at the stage boundaries (development timeline: tools/bench_configs.py c3 --timeline)."""
        pf = self.prepared
        caller = torch.cuda.current_stream()
        main = self._crit
        main.wait_stream(caller)

        def mark(name, stream=None):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream or main)
                marks[name] = ev
        with torch.cuda.stream(main):
            mark("start")
            if not self.empty:
                pf.entity(False); pf.meshlet_test(False, self.mx_early.local_masks)
            else:
                pf.early_dispatch[:4].zero_()
            mark("early tested")
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.mx_early.exchange(pf.early_dispatch)
                mark("early entries on rank 0", self._side)
            if self.rank == 0:
                pf.hiz()
            mark("hiz built")
            if self.pyr_bcast is not None:
                self.pyr_bcast.broadcast()
            else:
                broadcast_pyramid(self.vstate.depth_pyramid.texels, src=0)
            mark("pyramid broadcast")
            # Rank 0 emits the early list (C3: 229 MB of stores) only after the pyramid has left: emitted beside the Hi-Z build
            # it stretched that build from ~50 to 126 us on 8 GPUs — on the critical path of every rank — while beside rank
            # 0's own (lighter) late test it only delays rank 0 (profiles/r2_c3_timeline_n8.txt)
            if self.rank == 0:
                sent = torch.cuda.Event()
                sent.record(main)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(sent)
                    self.mx_early.expand(pf.sb_early, self.gathered_early, self.total_dcap)
            mark("early list emitted", self._side)
            if not self.empty:
                pf.entity(True); pf.meshlet_test(True, self.mx_late.local_masks)
            else:
                pf.late_dispatch[:4].zero_()
            mark("late tested")
            self.mx_late.exchange(pf.late_dispatch)
            mark("late entries on rank 0")
            if self.rank == 0:
                self.mx_late.expand(pf.sb_late, self.gathered_late, self.total_dcap)
            mark("late list emitted")
            main.wait_stream(self._side)
            mark("end")
        caller.wait_stream(main)

    def clear_gathered(self):
        if self.rank == 0:
            self.gathered_early.zero_(); self.gathered_late.zero_()
        torch.cuda.synchronize()

    def gathered_lists(self):
        """The two assembled MeshletDrawCommandBuffers (rank 0) after step_best()."""
        return self.gathered_early, self.gathered_late

    def close(self):
        for name in ("peer_early", "peer_late", "mx_early", "mx_late"):
            if hasattr(self, name):
                getattr(self, name).close()
                delattr(self, name)
        if getattr(self, "pyr_bcast", None) is not None:      # the pyramid lives in its memory: nothing may use the view afterwards
            torch.cuda.synchronize(self.context.device)
            dist.barrier()
            self.pyr_bcast.close()
            self.pyr_bcast = None
