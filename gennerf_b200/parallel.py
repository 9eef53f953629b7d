"""Multi-GPU partitioning of the lift-and-query path.

One process per GPU (torchrun), torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The path shards naturally, so collectives appear only where the data really
has to move:

  features  the 2D CNN ran on T/N frames per rank (or on one rank) -> ONE all-gather (or broadcast) of a flat
            channels-last buffer (FrameBuffer), so that every rank can lift
  lift      voxels are independent -> either every rank lifts the whole grid itself (0.4 ms at config 4: cheaper than
            moving the 805 MB volume), or contiguous x-slabs per rank + ONE in-place all-gather for grids where
            the lift is the larger term
  triplane  points split across ranks -> partial sums (fp32) + counts (int32) -> all-reduce,
            divide locally (counts stay exact; fp32 sums are order-dependent as in atomic mode)
  query     points are independent -> contiguous ranges of Q per rank, no collective
            (optionally gathered to every rank)

The compute backend is injected (`backend`): gennerf_b200.ops on the GPU; the CPU tests pass
an oracle-backed stand-in so the partition / collective logic is covered with world_size 2
on gloo.  Nothing here falls back to the CPU on its own.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [begin, end) of n items for `rank`: sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _ws(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def lift_sharded(backend, voxel_dim, voxel_size, origin, projections, features, group=None, gather=True):
    """Every rank holds the frame features (broadcast them first if only one rank ran the CNN,
    see broadcast_features) and lifts the x-slab shard_range(nx) of the grid; with gather=True the
    slabs are all-gathered so every rank ends up with the complete (volume, count, valid).

    Returns volume (B,C,nx,ny,nz) [channels-last storage], count (B,nx,ny,nz) int32,
    valid (B,1,nx,ny,nz) bool -- bit-identical to the single-GPU result (voxels are independent).
    """
    rank, world = _ws(group)
    nx, ny, nz = (int(d) for d in voxel_dim)
    B, C = features[0].shape[0], features[0].shape[1]
    dev = features[0].device
    x0, x1 = shard_range(nx, rank, world)
    store = torch.empty((B, nx, ny, nz, C), device=dev, dtype=torch.float32)
    count = torch.empty((B, nx, ny, nz), device=dev, dtype=torch.int32)
    valid = torch.empty((B, 1, nx, ny, nz), device=dev, dtype=torch.bool)
    volume = store.permute(0, 4, 1, 2, 3)
    backend.backproject_frames(voxel_dim, voxel_size, origin, projections, features, out=(volume, count, valid),
                               accumulate=False, x_range=(x0, x1))
    if gather and world > 1:
        for b in range(B):                       # a slab is one contiguous block per scene
            _all_gather_slabs(store[b], nx, world, group)
            _all_gather_slabs(count[b], nx, world, group)
            _all_gather_slabs(valid[b, 0].view(torch.uint8), nx, world, group)
    return volume, count, valid


def _in_place_ok(group):
    """NCCL's all-gather is in place when the send buffer is the rank's own slot of the receive buffer
    (sendbuff == recvbuff + rank * count); gloo is given a private copy."""
    try:
        return dist.get_backend(group) == "nccl"
    except Exception:
        return False


def _all_gather_slabs(t, nx, world, group):
    """In-place all-gather of the x-slabs of t (nx, ...): rank r owns rows shard_range(nx, r)."""
    rank, _ = _ws(group)
    bounds = [shard_range(nx, r, world) for r in range(world)]
    if nx % world == 0:
        x0, x1 = bounds[rank]
        mine = t[x0:x1]                                                     # equal slabs: one fused collective, and on
        dist.all_gather_into_tensor(t, mine if _in_place_ok(group) else mine.clone(), group=group)   # NCCL no staging copy
        return
    # ragged slabs: pad to the largest, gather, copy back (works on every backend)
    parts = _all_gather_ragged(t[bounds[rank][0]:bounds[rank][1]], [b - a for a, b in bounds], group)
    for r, (a, b) in enumerate(bounds):
        if r != rank:
            t[a:b] = parts[r]


def _all_gather_ragged(local, sizes, group):
    """all-gather of tensors whose dim 0 differs per rank (sizes[r]); returns the list of pieces."""
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[: local.shape[0]] = local
    buf = torch.empty((len(sizes) * mx,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return [buf[r * mx: r * mx + n] for r, n in enumerate(sizes)]


def broadcast_features(features, src=0, group=None):
    """Frame features live on the rank that ran the 2D CNN: broadcast them (NCCL over NVLink).  A list of separate
    tensors costs one collective per frame; hand over ONE flat buffer (FrameBuffer.flat) for a single call."""
    rank, world = _ws(group)
    if world > 1:
        if torch.is_tensor(features):
            dist.broadcast(features, src=src, group=group)
        else:
            for f in features:
                dist.broadcast(f, src=src, group=group)
    return features


class FrameBuffer:
    """All T frames' feature maps of one scene batch in ONE channels-last buffer (T,B,H,W,C), so that moving them
    between GPUs is a single collective on 1.26 GB (config 4) instead of one per frame.

    `frames` are the T logical (B,C,H,W) views the lift kernel consumes in place (channels_last memory format);
    `owned` = this rank's frame range shard_range(T): the frames whose 2D CNN ran here."""

    def __init__(self, T, B, C, H, W, device, group=None):
        self.group = group
        self.rank, self.world = _ws(group)
        self.T = T
        self.flat = torch.empty((T, B, H, W, C), device=device, dtype=torch.float32)
        self.frames = [self.flat[t].permute(0, 3, 1, 2) for t in range(T)]
        self.owned = shard_range(T, self.rank, self.world)

    def all_gather(self):
        """Every rank has filled its `owned` frames: afterwards every rank holds all T.  Each GPU sends its T/N frames to
        all peers at once through NVSwitch (N concurrent broadcasts): per-GPU traffic (N-1)/N of the buffer in each
        direction, instead of one source pushing the whole buffer."""
        if self.world == 1:
            return self
        t0, t1 = self.owned
        if self.T % self.world == 0:
            mine = self.flat[t0:t1]
            dist.all_gather_into_tensor(self.flat, mine if _in_place_ok(self.group) else mine.clone(), group=self.group)
        else:                                            # ragged ownership: one broadcast per owner
            for r in range(self.world):
                a, b = shard_range(self.T, r, self.world)
                if b > a:
                    dist.broadcast(self.flat[a:b], src=dist.get_global_rank(self.group, r) if self.group is not None else r,
                                   group=self.group)
        return self

    def broadcast(self, src=0):
        """The rank `src` holds all T frames (it ran the CNN alone): one broadcast of the flat buffer."""
        if self.world > 1:
            dist.broadcast(self.flat, src=src, group=self.group)
        return self


class P2PFrameBuffer:
    """FrameBuffer whose exchange does not use an SM: `depth` flat buffers in symmetric memory (every rank maps every peer's
    buffer, torch.distributed._symmetric_memory); after a device-side barrier on the signal pads each rank PULLS the other
    ranks' frame slices over NVLink with plain device-to-device copies, i.e. on the copy engines, on its own copy stream.
    The kernels of the scene being processed (the persistent tcgen05 decoder fills every SM for the whole query phase: an NCCL
    kernel launched beside it would wait for a free SM) therefore run undisturbed while the NEXT scene's frames arrive:

        write(k) own frames into slots[k] -> exchange(k) [barrier on the current stream, pulls on the copy stream]
        ... kernels of the current scene ...
        wait(k) -> lift from slots[k].frames

    Buffer reuse needs no second barrier: a rank rewrites its slice of slot k only after a later exchange() barrier, which every
    peer reaches only after its own wait(k), i.e. after its pulls from slot k have completed."""

    def __init__(self, T, B, C, H, W, device, group=None, depth=2, barrier_timeout_ms=60000):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = _ws(group)
        if T % self.world:
            raise ValueError("P2PFrameBuffer: the frames must split evenly over the ranks")
        self.T, self.per = T, T // self.world
        self.owned = shard_range(T, self.rank, self.world)
        self.timeout = int(barrier_timeout_ms)
        self.slots = []
        for _ in range(depth):
            flat = symm.empty((T, B, H, W, C), dtype=torch.float32, device=device)
            hdl = symm.rendezvous(flat, self.group)
            peers = [flat if r == self.rank else hdl.get_buffer(r, tuple(flat.shape), flat.dtype) for r in range(self.world)]
            self.slots.append({"flat": flat, "hdl": hdl, "peers": peers, "done": torch.cuda.Event(),
                               "frames": [flat[t].permute(0, 3, 1, 2) for t in range(T)]})
        self.copy_stream = torch.cuda.Stream(device)

    def own(self, k):
        """This rank's slice of slot k (channels-last (T/N, B, H, W, C)): where its frames are written before exchange(k)."""
        return self.slots[k]["flat"][self.owned[0]:self.owned[1]]

    def frames(self, k):
        return self.slots[k]["frames"]

    def exchange(self, k):
        """Every rank has written own(k) on its current stream.  Returns at once; wait(k) orders later work behind the pulls."""
        s = self.slots[k]
        if self.world > 1:
            s["hdl"].barrier(channel=0, timeout_ms=self.timeout)          # on the current stream: all slices of slot k are written
        go = torch.cuda.Event()
        go.record()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(go)
            for i in range(1, self.world):                                 # staggered sources: every GPU serves one reader at a time
                r = (self.rank + i) % self.world
                a, b = r * self.per, (r + 1) * self.per
                s["flat"][a:b].copy_(s["peers"][r][a:b], non_blocking=True)
            s["done"].record()

    def wait(self, k):
        torch.cuda.current_stream().wait_event(self.slots[k]["done"])


def scatter_planes_sharded(backend, p, c, reso, padding=0.1, group=None):
    """p (B,N,3), c (B,N,C_p): every rank scatters ITS points (the caller passes the rank's
    shard); partial sums and counts are all-reduced, then divided locally.
    Returns planes (3,B,C_p,R,R) [channels-last storage], count (3,B,R,R) int32 on every rank."""
    rank, world = _ws(group)
    sums, count = backend.scatter_mean_planes(p, c, reso, padding, "sum")
    if world > 1:
        store = sums.permute(0, 1, 3, 4, 2)
        dist.all_reduce(store, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    return backend.scatter_finalize(sums, count), count


def query_sharded(query_fn, xyz, group=None, gather=False):
    """xyz (B,Q,3) replicated on every rank (or only the rank's range if pre-sharded=False is not
    needed): each rank answers the contiguous range shard_range(Q) with `query_fn(xyz_shard)` ->
    tensor (B,q,...) and, with gather=True, the ranges are concatenated on every rank."""
    rank, world = _ws(group)
    Q = xyz.shape[1]
    q0, q1 = shard_range(Q, rank, world)
    out = query_fn(xyz[:, q0:q1].contiguous())
    if not gather or world == 1:
        return out, (q0, q1)
    sizes = [shard_range(Q, r, world)[1] - shard_range(Q, r, world)[0] for r in range(world)]
    parts = _all_gather_ragged(out.transpose(0, 1).contiguous(), sizes, group)       # ranges along dim 0
    return torch.cat(parts, dim=0).transpose(0, 1).contiguous(), (0, Q)
