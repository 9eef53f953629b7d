"""One scene sharded over N GPUs (BASELINE config 4 shape, scaled by --scale): x-slab lift +
NCCL all-gather of the volume, query ranges per rank, fused tcgen05 decoder.  Verifies the
sharded lift against the single-GPU lift bit-for-bit, then reports device-timed (max over
ranks) lift / gather / query times.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [--cfg cfg4]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, parallel, synthetic as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg4")
ap.add_argument("--queries", type=int, default=0, help="total queries (default: the config's)")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")        # NCCL's version banner / debug lines: not on stdout
    dist.init_process_group("nccl", device_id=dev)
VS, C = 0.04, 32
wl = S.WORKLOADS[args.cfg]
Q = args.queries or wl["Q"]
g = S.gen(1004)
origin = torch.tensor([0, 0, 0]).view(1, 3)
T = wl["T"]
P = S.projections(T, wl["H"], wl["W"], wl["voxel_dim"], VS, g).unsqueeze(0)
gd = torch.Generator(device=dev)
gd.manual_seed(5)
# frame features exist on rank 0 (it "ran the CNN"), channels-last; the others receive the broadcast
feats = [torch.randn(1, wl["H"], wl["W"], C, device=dev, generator=gd).permute(0, 3, 1, 2) if rank == 0
         else torch.empty(1, wl["H"], wl["W"], C, device=dev).permute(0, 3, 1, 2) for _ in range(T)]
w, hw, hb = S.decoder_weights(g, C, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
dw.pack("fp16")
q0, q1 = parallel.shard_range(Q, rank, world)
gq = torch.Generator(device=dev)
gq.manual_seed(100 + rank)
ext = torch.tensor([d * VS for d in wl["voxel_dim"]], device=dev)
xyz = ((torch.rand(1, q1 - q0, 3, device=dev, generator=gq) * 1.1 - 0.05) * ext).contiguous()


def ev():
    return torch.cuda.Event(enable_timing=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


times = []
for it in range(args.steps + 2):
    barrier()
    e = [ev() for _ in range(5)]
    e[0].record()
    parallel.broadcast_features(feats, src=0)
    e[1].record()
    vol, cnt, valid = parallel.lift_sharded(ops, wl["voxel_dim"], VS, origin, P, feats, gather=False)
    e[2].record()
    if world > 1:
        nx = wl["voxel_dim"][0]
        store = vol.permute(0, 2, 3, 4, 1)
        parallel._all_gather_slabs(store[0], nx, world, None)
        parallel._all_gather_slabs(cnt[0], nx, world, None)
        parallel._all_gather_slabs(valid[0, 0].view(torch.uint8), nx, world, None)
    e[3].record()
    out, tsdf, _ = ops.query_fused(dw, xyz, volume=vol, voxel_size=VS, origin=origin, want_feat=False)
    e[4].record()
    torch.cuda.synchronize()
    if it >= 2:
        times.append([e[i].elapsed_time(e[i + 1]) for i in range(4)])
t = torch.tensor(times, device=dev, dtype=torch.float64).mean(0)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# correctness: the gathered volume equals the single-GPU lift
vol1, cnt1, valid1 = ops.backproject_frames(wl["voxel_dim"], VS, origin, P, feats)
ok = torch.equal(vol1, vol) and torch.equal(cnt1, cnt) and torch.equal(valid1, valid)
okt = torch.tensor([int(ok)], device=dev)
if world > 1:
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    V = cnt.numel()
    bc, lift, gather, query = t.tolist()
    print(json.dumps({"cfg": args.cfg, "n_gpus": world, "queries_total": Q, "sharded_lift_equals_single_gpu": bool(okt.item()),
                      "ms": {"broadcast_features": bc, "lift_slab": lift, "all_gather_volume": gather, "query_range": query},
                      "voxel_frames_per_s": V * T / ((lift + gather) * 1e-3), "tsdf_points_per_s": Q / (query * 1e-3),
                      "bytes": {"features": T * wl["H"] * wl["W"] * C * 4, "volume": V * C * 4}}))
if world > 1:
    dist.destroy_process_group()
