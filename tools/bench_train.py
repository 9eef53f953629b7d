"""BASELINE config 5: one training step (forward + L1 TSDF loss + backward + Adam) of the drop-in GenNerf on the B200
path, one scene per GPU, data-parallel over N GPUs with an NCCL all-reduce of the gradients.

Per scene: T = 8 frames of 480x640x32ch feature maps (requires_grad: they stand for the CNN output),
160x160x64 grid @ 4 cm, triplane branch from the depth maps (unprojection + FPS 512 points per frame + PointNet ->
3 x 128^2 x 32 planes), 2900 query points per frame, default MLP (d_hidden 512, 5 blocks).

  python tools/bench_train.py [--steps K] [--warmup W]                 # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_train.py

Prints one JSON line (rank 0): scenes/s over all ranks, ms per step (max over ranks), the per-phase breakdown of rank 0.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import synthetic as S  # noqa: E402
from gennerf_b200.dropin import GenNerf  # noqa: E402


class Attr(dict):
    __getattr__ = dict.get


def attr(d):
    return Attr({k: attr(v) if isinstance(d[k], dict) else v for k, v in d.items()}) if isinstance(d, dict) else d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--matmul-precision", default="high", choices=["highest", "high"],
                    help="torch.set_float32_matmul_precision for the PyTorch part (the MLP's dgrad/wgrad): 'high' (TF32) is what "
                         "the reference's training sets (src/utils/utils.py:48); 'highest' = fp32 SIMT GEMMs")
    ap.add_argument("--train-precision", default="fp16", choices=["fp16", "fp32"],
                    help="MLP of the training step: fp16 = tcgen05 forward kernel with saved activations + backward built on them "
                         "(gennerf_b200/train_decode.py); fp32 = nn.Linear under autograd (ResnetFC.forward_torch)")
    ap.add_argument("--adam", default="foreach", choices=["foreach", "fused"],
                    help="torch.optim.Adam implementation (the optimiser is outside the path; the reference's Hydra config instantiates "
                         "torch.optim.Adam with PyTorch's default, foreach): 'fused' = one kernel for all parameters, ~0.7 ms less host time")
    ap.add_argument("--graph-pointnet", action="store_true",
                    help="experiment: the PointNet encoder (forward + backward, ~150 launches of a few microseconds) as two CUDA graphs "
                         "(torch.cuda.make_graphed_callables).  The ops capture as they are, and the host's issue time drops from 6.2 to "
                         "4.9 ms per step, but the replayed graphs take longer on the GPU than the stream launches they replace "
                         "(step 6.2 -> 7.3 ms): off by default")
    args = ap.parse_args()
    torch.set_float32_matmul_precision(args.matmul_precision)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # NCCL's version banner / debug lines: not on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    wl = S.WORKLOADS["cfg5"]
    VS, C, Cp = 0.04, 32, 32
    T, H, W, vd, Q, R = wl["T"], wl["H"], wl["W"], wl["voxel_dim"], wl["Q"], wl["R"]
    cfg = attr({
        "voxel_size": VS, "voxel_dim_train": list(vd), "voxel_dim_val": list(vd),
        "encoder": {"use_spatial": True, "spatial": {"num_layers": 1, "latent_size": C}, "use_pointnet": True, "use_auxiliary": False,
                    "pointnet": {"num_sparse_points": 512, "c_dim": Cp, "dim": 3, "padding": 0.1, "hidden_dim": 32,
                                 "scatter_type": "max", "plane_type": ["xz", "xy", "yz"], "plane_resolution": R,
                                 "n_blocks": 5, "unet": False, "unet_kwargs": None, "sample_mode": "bilinear"},
                    "plane_merger": {"strategy": "average", "alpha": 0.1}},
        "mlp": {"d_out_sem": 32, "d_out_geo": 32, "n_blocks": 5, "d_hidden": 512, "combine_layer": 1000,
                "combine_type": "average", "beta": 0.0, "use_spade": False, "use_layer_norm": False, "alpha": 1.0},
        "use_code": True, "code": {"num_freqs": 2, "freq_factor": 0.5, "include_input": True},
    })
    torch.manual_seed(7)                                        # same initial weights on every rank
    model = GenNerf(cfg, precision="fp16", fused=True, train_precision=args.train_precision).to(dev).train()
    if args.graph_pointnet:
        # static shapes, device-resident inputs, no host decisions inside: the module and the kernels under it capture as they are
        sample = (torch.rand(1, T * cfg.encoder.pointnet.num_sparse_points, 3, device=dev) * 4.0,)
        model.pointnet = torch.cuda.make_graphed_callables(model.pointnet, sample)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, **({"fused": True} if args.adam == "fused" else {}))
    g = S.gen(5000 + rank)                                      # a different scene per rank
    P = S.projections(T, H, W, vd, VS, g).unsqueeze(0)
    feats = torch.randn(1, T, C, H, W, generator=g).to(dev)
    depth = S.surface_depth_maps(T, H, W, g, mean=1.5, holes=False).unsqueeze(0).to(dev)
    xyz = S.query_points(Q, vd, VS, g).to(dev)
    target = (torch.rand(1, Q, 1, generator=g) * 2 - 1).to(dev)
    sizes = [p.numel() for p in params]

    ev = lambda: torch.cuda.Event(enable_timing=True)           # noqa: E731

    def step(marks=None):
        def mark(name):
            if marks is not None:
                e = ev()
                e.record()
                marks.append((name, e))
        mark("start")
        f = feats.detach().requires_grad_(True)
        model.initialize_volume()
        model.encode(P, f, depth, "train")
        mark("encode")
        out = model(xyz)
        loss = (out["tsdf"] - target).abs().mean()
        mark("query")
        opt.zero_grad(set_to_none=True)
        loss.backward()
        mark("backward")
        if world > 1:                                           # data-parallel: average the gradients over the scenes
            # one flat bucket: a concatenation (one kernel), ONE NCCL all-reduce, one multi-tensor copy back
            for p_ in params:
                if p_.grad is None:
                    p_.grad = torch.zeros_like(p_)
            bucket = torch.cat([p_.grad.reshape(-1) for p_ in params])
            dist.all_reduce(bucket)
            bucket.div_(world)
            torch._foreach_copy_([p_.grad for p_ in params], [c.view_as(p_) for c, p_ in zip(bucket.split(sizes), params)])
            mark("allreduce")
        opt.step()
        mark("adam")
        return loss, f.grad

    for _ in range(max(args.warmup, 3)):
        loss, gf = step()
        if os.environ.get("TRAIN_TRACE"):
            print(f"warm-up step: loss {float(loss):.6f} grad_feature_norm {float(gf.norm()):.6e}", file=sys.stderr)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = ev(), ev()
    a.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        loss, gf = step()
    host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps     # host time to ISSUE a step (no sync inside): >= ms means host-bound
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    marks = []
    step(marks)
    torch.cuda.synchronize()
    phases = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
    if os.environ.get("TRAIN_PROFILE") and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70), file=sys.stderr)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "metric": "training_scenes_per_s", "value": world / (t.item() * 1e-3), "unit": "scenes/s", "n_gpus": world,
            "steps": args.steps, "ms_per_step": t.item(), "host_issue_ms_per_step_rank0": host_ms, "scaling": "weak",
            "config": {"workload": "BASELINE config 5: fwd + L1 TSDF loss + bwd + Adam, one scene per GPU: 8 frames 480x640x32ch, "
                                   "160x160x64 grid, FPS 512 pts/frame -> 3x128^2x32 planes, 23200 queries, MLP 512x5",
                       "parallelism": f"dp{world} (NCCL all-reduce of {sum(sizes)} gradient elements)",
                       "float32_matmul_precision": args.matmul_precision, "train_precision": args.train_precision, "adam": args.adam,
                       "pointnet_cuda_graph": bool(args.graph_pointnet)},
            "phases_ms_rank0": phases, "loss": float(loss), "grad_feature_norm": float(gf.norm())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
