"""Multi-process (world_size 2, gloo, CPU) coverage of the sharding / collective logic in
gennerf_b200.parallel.  The per-rank compute is an oracle-backed stand-in with the same call
signatures as gennerf_b200.ops (the kernels need a GPU; the partition logic does not)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gennerf_b200 import parallel
from gennerf_b200 import synthetic as S
from oracle import gennerf_oracle as O

ORIGIN = torch.tensor([0, 0, 0]).view(1, 3)
VS = 0.04


class OracleBackend:
    """ops-compatible stand-in running the CPU oracle (test infrastructure)."""

    @staticmethod
    def backproject_frames(voxel_dim, voxel_size, origin, projections, features, out=None, accumulate=False, x_range=None):
        vol, valid, cnt = O.encode_volume(voxel_dim, voxel_size, origin, projections, features)
        volume, count, val = out
        x0, x1 = x_range
        volume[:, :, x0:x1] = vol[:, :, x0:x1]
        count[:, x0:x1] = cnt[:, x0:x1]
        val[:, :, x0:x1] = valid[:, :, x0:x1]
        return volume, count, val

    @staticmethod
    def scatter_mean_planes(p, c, reso, padding, mode):
        assert mode == "sum"
        B, N, Cp = c.shape
        sums = torch.zeros(3, B, reso, reso, Cp)
        count = torch.zeros(3, B, reso, reso, dtype=torch.int32)
        for k, name in enumerate(O.PLANES):
            xy = O.normalize_coordinate(p.clone(), padding, name)
            idx = O.coordinate2index(xy, reso)
            s = torch.zeros(B, Cp, reso * reso).scatter_add_(2, idx.expand(B, Cp, N), c.permute(0, 2, 1))
            n = torch.zeros(B, 1, reso * reso).scatter_add_(2, idx, torch.ones(B, 1, N))
            sums[k] = s.view(B, Cp, reso, reso).permute(0, 2, 3, 1)
            count[k] = n.view(B, reso, reso).to(torch.int32)
        return sums.permute(0, 1, 4, 2, 3), count

    @staticmethod
    def scatter_finalize(planes, count):
        planes /= count.clamp_min(1).unsqueeze(2).float()
        return planes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nx, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        g = S.gen(7)
        vd = (nx, 10, 6)
        T = 3
        P = S.projections(T, 24, 32, vd, VS, g, pull_back=0.8).unsqueeze(0)
        feats = S.frame_features(T, 4, 24, 32, g)
        # features "live on rank 0": other ranks start from garbage and receive the broadcast
        mine = [f.clone() if rank == 0 else torch.full_like(f, float("nan")) for f in feats]
        parallel.broadcast_features(mine, src=0)
        vol, cnt, valid = parallel.lift_sharded(OracleBackend, vd, VS, ORIGIN, P, mine)
        vol_o, valid_o, cnt_o = O.encode_volume(vd, VS, ORIGIN, P, feats)
        assert torch.equal(vol, vol_o) and torch.equal(cnt, cnt_o) and torch.equal(valid, valid_o)

        # frames owned per rank (the CNN ran on T/N frames each) -> one flat all-gather (equal and ragged ownership)
        for Tn in (4, 3):
            fr = S.frame_features(Tn, 4, 6, 8, S.gen(8))
            fb = parallel.FrameBuffer(Tn, 1, 4, 6, 8, "cpu")
            fb.flat.fill_(float("nan"))
            for t in range(*fb.owned):
                fb.frames[t].copy_(fr[t])
            fb.all_gather()
            for t in range(Tn):
                assert torch.equal(fb.frames[t], fr[t]) and fb.frames[t].shape == fr[t].shape
            fb2 = parallel.FrameBuffer(Tn, 1, 4, 6, 8, "cpu")
            if rank == 0:
                for t in range(Tn):
                    fb2.frames[t].copy_(fr[t])
            fb2.broadcast(src=0)
            assert all(torch.equal(fb2.frames[t], fr[t]) for t in range(Tn))

        # triplane: each rank scatters its share of the points; the all-reduced result equals the full scatter
        N, Cp, R = 1001, 4, 8
        p = S.plane_points(N, g, "unit")
        c = torch.randn(1, N, Cp, generator=g)
        a, b = parallel.shard_range(N, rank, world)
        planes, count = parallel.scatter_planes_sharded(OracleBackend, p[:, a:b], c[:, a:b], R, 0.1)
        for k, name in enumerate(O.PLANES):
            ref, cnt_ref = O.generate_plane_features(p, c, name, R, 0.1, return_count=True)
            assert torch.equal(count[k], cnt_ref), "counts are exact integers whatever the partition"
            assert torch.allclose(planes[k], ref, rtol=1e-5, atol=1e-6)

        # queries: contiguous ranges, gathered
        xyz = S.query_points(103, vd, VS, g)
        out, rng = parallel.query_sharded(lambda x: x.sum(-1, keepdim=True) * 2.0, xyz, gather=True)
        assert rng == (0, 103) and torch.equal(out, xyz.sum(-1, keepdim=True) * 2.0)
        part, rng = parallel.query_sharded(lambda x: x[..., :1], xyz, gather=False)
        assert rng == parallel.shard_range(103, rank, world) and torch.equal(part, xyz[:, rng[0]:rng[1], :1])
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx", [12, 13])          # equal slabs (fused all-gather) and ragged slabs
def test_sharded_lift_scatter_query_world2(tmp_path, nx):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nx, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 96, 1 << 20):
        for world in (1, 2, 3, 8):
            pieces = [parallel.shard_range(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1
