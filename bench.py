#!/usr/bin/env python
"""Benchmark of the lift-and-query hot path (BASELINE.json metric: voxel.frames/s back-projection; TSDF query points/s).

One "step" = ONE SCENE of BASELINE config 4 through the whole path, on N GPUs of one box (strong scaling: the scene is
fixed, the GPUs share it as north_star partitions it):
  features  32 frames of 480x640x32 feature maps in the reference's NCHW layout, T/N frames per rank (the rank that ran
            the 2D CNN on them): NCHW->NHWC into the rank's slot of ONE flat buffer, then ONE NCCL all-gather (N > 1)
  lift      every rank lifts the whole 256x256x96 grid itself (cheaper than moving the 805 MB volume; --lift slab
            measures the x-slab + all-gather alternative)
  planes    3 x 256^2 x 32 triplanes: each rank scatters ITS frames' points (32 x 512 in total), partial sums and counts
            are all-reduced with NCCL, divided locally
  query     16 Mi TSDF queries in contiguous ranges of Q/N per rank: trilinear + 3 x bilinear sampler fused into the
            tcgen05 ResNet-MLP decoder; no collective on this path
`value` = 16 Mi / (time of the whole step), device-timed, max over ranks; inputs resident in HBM (1.26 GB of features and
an 805 MB volume: larger than the 126 MB L2, so no flush is needed between steps).  `e2e` = the same scene through the
drop-in `GenNerf.shard_scene / encode / forward` with pinned HOST buffers: H2D of the rank's frames and query range and
D2H of its TSDF range inside the timed region.  `--impl reference` times the reference's CPU algorithm (oracle port: the
same ATen CPU kernels the reference calls) on the host cores on a bounded sample of the same scene.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gennerf_b200 import synthetic as S   # noqa: E402

VS = 0.04
C_FEAT, C_PLANE, R_PLANE, PTS_PER_FRAME = 32, 32, 256, 512
MLP = dict(d_hidden=512, n_blocks=5, d_out=64, d_geo=32, num_freqs=2, freq_factor=0.5)
D_CODE = 3 + 6 * MLP["num_freqs"]
WL = S.WORKLOADS["cfg4"]
SEED = 1004


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def flops_per_query(d_feat, d_code, d_hidden, n_blocks, d_out, d_geo):
    return 2 * (d_feat * d_hidden + n_blocks * (d_code * d_hidden + 2 * d_hidden * d_hidden) + d_hidden * d_out + d_geo)


def lift_bytes(T, C, H, W, V, n_valid):
    """every input element once (or only the gathered ones if fewer), every output once."""
    return min(T * C * H * W * 4, n_valid * C * 4) + V * C * 4 + V * 5 + T * 48


def config_dict():
    return {"workload": "BASELINE config 4 as ONE scene: combined volume + triplane, 32 synthetic frames 480x640 x 32 ch -> "
                        "256x256x96 grid @4cm + 3x256^2x32 planes from 32x512 points, 16 Mi TSDF queries through the fused "
                        "sampler + ResNet-MLP decoder (d_hidden 512, 5 blocks, d_out 32+32); frames, points and queries sharded "
                        "over the GPUs",
            "frames": WL["T"], "image": [WL["H"], WL["W"]], "channels": C_FEAT, "grid": list(WL["voxel_dim"]),
            "planes": [3, R_PLANE, R_PLANE, C_PLANE], "plane_points": WL["T"] * PTS_PER_FRAME, "queries": WL["Q"],
            "l2": "inputs larger than L2 (1.26 GB features, 805 MB volume, 201 MB queries vs 126 MB L2): no flush between steps"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        hi = [s for s in sm if s >= 0.5 * mx] or sm
        return {"sm_mhz": hi[len(hi) // 2] if hi else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# synthetic scene (identical bits on every rank and in both arms: everything from one seeded CPU generator)
# ------------------------------------------------------------------------------------------------
def scene_small_parts():
    """Everything but the frames and the queries (cheap): projections, plane points + features, decoder weights."""
    g = S.gen(SEED)
    P = S.projections(WL["T"], WL["H"], WL["W"], WL["voxel_dim"], VS, g).unsqueeze(0)
    N = WL["T"] * PTS_PER_FRAME
    pts = S.plane_points(N, g, "metric", voxel_dim=WL["voxel_dim"])        # metres, as GenNerf.encode really feeds them (trap T6)
    cpt = torch.randn(1, N, C_PLANE, generator=g)
    w, hw, hb = S.decoder_weights(g, C_FEAT + C_PLANE, D_CODE, MLP["d_hidden"], MLP["n_blocks"], MLP["d_out"], MLP["d_geo"])
    return P, pts, cpt, (w, hw, hb)


def frame(t):
    """Frame t's (1,C,H,W) feature map, NCHW like the reference's CNN output; per-frame seed so that a rank can make its own."""
    return torch.randn(1, C_FEAT, WL["H"], WL["W"], generator=S.gen(SEED * 1000 + t))


def queries(q0, q1, chunk=1 << 20):
    """Queries [q0, q1) of the scene's Q: generated in 1 Mi blocks with per-block seeds (any rank can make any range)."""
    out = []
    for b0 in range(q0 - q0 % chunk, q1, chunk):
        blk = S.query_points(chunk, WL["voxel_dim"], VS, S.gen(SEED * 7919 + b0 // chunk))
        out.append(blk[:, max(q0 - b0, 0):min(q1 - b0, chunk)])
    return torch.cat(out, dim=1).contiguous()


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port) on the host cores, bounded sample of the same scene
# ------------------------------------------------------------------------------------------------
def cpu_scene_time(n_frames, n_queries, P, pts, cpt, weights):
    """Seconds of the oracle for: lift of `n_frames` frames into the full grid, the full triplane scatter, and
    `n_queries` queries in the reference's 10 000-point chunks (model.py:769-777, volume re-normalised per chunk)."""
    from oracle import gennerf_oracle as O
    w, hw, hb = weights
    origin = torch.tensor([0, 0, 0]).view(1, 3)
    with torch.no_grad():
        feats = [frame(t) for t in range(n_frames)]
        t0 = time.perf_counter()
        vol, valid, _ = O.encode_volume(WL["voxel_dim"], VS, origin, P[:, :n_frames], feats)
        t1 = time.perf_counter()
        planes = {k: O.generate_plane_features(pts, cpt, k, R_PLANE, 0.1) for k in O.PLANES}
        t2 = time.perf_counter()
        xyz = queries(0, n_queries)
        t3 = time.perf_counter()
        for q0 in range(0, n_queries, 10000):
            O.gennerf_forward(xyz[:, q0:q0 + 10000], w, hw, hb, volume=vol, valid=valid, planes=planes, voxel_size=VS, padding=0.1,
                              num_freqs=MLP["num_freqs"], freq_factor=MLP["freq_factor"], n_blocks=MLP["n_blocks"],
                              d_out_geo=MLP["d_geo"], d_out_sem=MLP["d_out"] - MLP["d_geo"])
        t4 = time.perf_counter()
    return t1 - t0, t2 - t1, t4 - t3


def cpu_estimate(n_frames, n_queries, parts):
    P, pts, cpt, weights = parts
    tl, tp, tq = cpu_scene_time(n_frames, n_queries, P, pts, cpt, weights)
    Q, T = WL["Q"], WL["T"]
    total = tl * (T / n_frames) + tp + tq * (Q / n_queries)
    V = WL["voxel_dim"][0] * WL["voxel_dim"][1] * WL["voxel_dim"][2]
    return total, {"lift_s_per_frame": tl / n_frames, "planes_s": tp, "query_s_per_10k": tq / (n_queries / 10000),
                   "lift_voxel_frames_per_s": V * n_frames / tl}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    parts = scene_small_parts()
    n_frames, n_queries = 2, 10000
    times, detail = [], None
    for i in range(args.warmup + args.steps):
        t, detail = cpu_estimate(n_frames, n_queries, parts)
        if i >= args.warmup:
            times.append(t)
    t = sum(times) / len(times)
    val = WL["Q"] / t
    sample = (f"per step: lift of {n_frames} of {WL['T']} frames into the full grid (scaled x{WL['T'] // n_frames}), the full triplane "
              f"scatter, {n_queries} of {WL['Q']} queries as one of the reference's 10k chunks incl. its per-chunk volume "
              f"re-normalisation (scaled x{WL['Q'] // n_queries})")
    line = {"impl": "reference", "metric": "tsdf_query_points_per_s", "value": val, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(),
            "cpu_baseline": {"value": val, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                             "detail": detail},
            "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def hbm_side_kernels(ops, dev, pk):
    """The HBM-bound kernels of the path timed alone on rank 0 (not part of `value`), BASELINE config 2 / 3 shapes:
    the lift (NCHW and channels-last input), the gather-only sampler on that volume with 1 Mi queries, config 3's
    triplane scatter.  Algorithmic bytes as SURVEY 8d; CUDA-graph replays, L2 flushed between replays, median of 10."""
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    origin = torch.tensor([0, 0, 0]).view(1, 3)

    def timed(fn):
        fn()
        g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        return sorted(ms)[len(ms) // 2]

    def roof(byt, ms):
        gbs = byt / (ms * 1e-3) / 1e9
        return {"ms": ms, "algorithmic_bytes": byt, "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"]}}

    out = {}
    wl = S.WORKLOADS["cfg2"]
    g = S.gen(1002)
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g).unsqueeze(0)
    feats = [f.to(dev) for f in S.frame_features(wl["T"], C_FEAT, wl["H"], wl["W"], g)]
    feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    xyz = S.query_points(1 << 20, wl["voxel_dim"], VS, g).to(dev)
    vol, cnt, _ = ops.backproject_frames(wl["voxel_dim"], VS, origin, P, feats_cl)
    V = cnt.numel()
    lb = lift_bytes(wl["T"], C_FEAT, wl["H"], wl["W"], V, int(cnt.sum().item()))
    out["lift_cfg2_nchw_input"] = dict(roof(lb, timed(lambda: ops.backproject_frames(wl["voxel_dim"], VS, origin, P, feats))),
                                       includes="NCHW->NHWC pass + lift kernel", voxel_frames=V * wl["T"])
    out["lift_cfg2_channels_last_input"] = dict(roof(lb, timed(lambda: ops.backproject_frames(wl["voxel_dim"], VS, origin, P, feats_cl))),
                                                includes="lift kernel only", voxel_frames=V * wl["T"])
    Q, C = xyz.shape[1], vol.shape[1]
    byt = Q * (12 + 4 * C) + min(vol.numel() * 4, 8 * Q * C * 4)
    for name, binned in (("sampler_binned", True), ("sampler_staged", False)):
        out[name] = dict(roof(byt, timed(lambda: ops.sample_features(xyz, volume=vol, voxel_size=VS, binned=binned))), queries=Q)
    g = S.gen(1003)
    for N in (4096, 614400):
        p = S.plane_points(N, g, "unit").to(dev)
        c = torch.randn(1, N, C_PLANE, generator=g).to(dev)
        byt = N * (12 + 4 * C_PLANE) + 3 * R_PLANE * R_PLANE * (4 * C_PLANE + 4)
        out[f"scatter_mean_planes_N{N}"] = roof(byt, timed(lambda: ops.scatter_mean_planes(p, c, R_PLANE, 0.1, "atomic")))
    del vol, feats, feats_cl
    # ... and the lift at the default yaml's 512 spatial channels (4 channel chunks of 128 per voxel brick), config-2 shape
    g = S.gen(1006)
    f512 = [torch.randn(1, wl["H"], wl["W"], 512, generator=g).to(dev).permute(0, 3, 1, 2) for _ in range(wl["T"])]
    lb512 = lift_bytes(wl["T"], 512, wl["H"], wl["W"], V, int(cnt.sum().item()))
    out["lift_cfg2_512_channels_last"] = dict(roof(lb512, timed(lambda: ops.backproject_frames(wl["voxel_dim"], VS, origin, P, f512))),
                                              includes="lift kernel only, C = 512", voxel_frames=V * wl["T"])
    del f512
    # The reference's DEFAULT Hydra config (configs/model/gen_nerf.yaml:43,56): 512 spatial + 32 plane channels = latent 544,
    # nine lin_in k-chunks streamed from the operand image.  1 Mi queries against a 512-channel volume on the config-2 grid.
    g = S.gen(1004)
    Cw = 512
    volw = (torch.randn(1, *wl["voxel_dim"], Cw, generator=g) * 0.2).to(dev).permute(0, 4, 1, 2, 3)
    plw = {k: (torch.randn(1, C_PLANE, R_PLANE, R_PLANE, generator=g) * 0.2).to(dev).contiguous(memory_format=torch.channels_last)
           for k in ops.PLANES}
    ww, hww, hbw = S.decoder_weights(S.gen(1005), Cw + C_PLANE, D_CODE, MLP["d_hidden"], MLP["n_blocks"], MLP["d_out"], MLP["d_geo"])
    dww = ops.DecoderWeights(ww, hww, hbw, n_blocks=MLP["n_blocks"], d_geo=MLP["d_geo"], num_freqs=MLP["num_freqs"], freq_factor=MLP["freq_factor"], device=dev)
    run = lambda: ops.query_fused(dww, xyz, volume=volw, planes=plw, voxel_size=VS, origin=origin, padding=0.1, want_feat=False)    # noqa: E731
    run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    m = sorted(ms)[len(ms) // 2]
    Hd = MLP["d_hidden"]
    fl = 2.0 * ((Cw + C_PLANE) * Hd + MLP["n_blocks"] * (D_CODE * Hd + 2 * Hd * Hd) + Hd * MLP["d_out"] + MLP["d_geo"]) * Q
    out["query_default_yaml_latent_544"] = {
        "ms": m, "queries": Q, "points_per_s": Q / (m * 1e-3), "includes": "sampler (512-channel volume + 32-channel planes) writing the 9-chunk "
        "operand image + tcgen05 decoder streaming it, 3 chunks of queries", "fp16_saturated": bool(dww.overflowed()),
        "roofline": {"bound": "tensor", "achieved": fl / (m * 1e-3) / 1e12, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                     "frac": fl / (m * 1e-3) / 1e12 / pk["bf16_burst"]}}
    return out


def gpu_eager_baseline(dev, P, pts, cpt, weights, n_frames=4, n_chunks=50):
    """Informational: the reference's own ATen CUDA path on this B200 (bmm / index_put_ / grid_sample / TF32 F.linear as
    src/utils/utils.py:48 sets it, 10 000-point chunks with .cpu() per chunk, model.py:769-777), bounded sample, scaled."""
    from oracle import eager_gpu as E
    w, hw, hb = weights
    old = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    try:
        origin = torch.tensor([0, 0, 0]).view(1, 3)
        wd = {k: v.to(dev) for k, v in w.items()}
        hwd, hbd = hw.to(dev), hb.to(dev)
        feats = [frame(t).to(dev) for t in range(n_frames)]
        Pd = P.to(dev)
        with torch.no_grad():
            E.encode_volume(WL["voxel_dim"], VS, origin, Pd[:, :1], feats[:1])       # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            vol, valid = E.encode_volume(WL["voxel_dim"], VS, origin, Pd[:, :n_frames], feats)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            planes = {k: E.generate_plane_features(pts.to(dev), cpt.to(dev), k, R_PLANE, 0.1) for k in E.PLANES}
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            xyz = queries(0, n_chunks * 10000).to(dev)
            E.predict_chunks(xyz[:, :20000], 10000, wd, hwd, hbd, vol, valid, planes, VS, 0.1, MLP["num_freqs"], MLP["freq_factor"],
                             MLP["n_blocks"], MLP["d_geo"])
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            E.predict_chunks(xyz, 10000, wd, hwd, hbd, vol, valid, planes, VS, 0.1, MLP["num_freqs"], MLP["freq_factor"],
                             MLP["n_blocks"], MLP["d_geo"])
            torch.cuda.synchronize()
            t4 = time.perf_counter()
        tl, tp, tq = t1 - t0, t2 - t1, t4 - t3
        Q, T = WL["Q"], WL["T"]
        total = tl * T / n_frames + tp + tq * Q / (n_chunks * 10000)
        V = WL["voxel_dim"][0] * WL["voxel_dim"][1] * WL["voxel_dim"][2]
        return {"value": Q / total, "unit": "points/s", "ms_per_step_scaled": total * 1e3,
                "what": "the reference's own PyTorch code path (ATen CUDA kernels, TF32 matmul, 10k-query chunks with a D2H copy "
                        "and a volume re-normalisation per chunk) on the same B200, one GPU",
                "sample": f"lift of {n_frames} of {T} frames (scaled), full triplane scatter, {n_chunks} of {Q // 10000} query chunks (scaled)",
                "lift_voxel_frames_per_s": V * n_frames / tl, "query_points_per_s": n_chunks * 10000 / tq}
    finally:
        torch.set_float32_matmul_precision(old)


def parity_check(ops, dev, P, feats_all, vol, planes, tsdf, xyz_h, weights, n_check=2000):
    """Bench-time parity (rank 0): (1) the lift of the scene's first two frames into the full grid == the CPU oracle,
    bit for bit; (2) `n_check` of the TSDFs this run produced, against the fp32 oracle answering the same points on the
    volume / planes this run built (copied to the host): |dTSDF| <= 1e-2."""
    from oracle import gennerf_oracle as O
    origin = torch.tensor([0, 0, 0]).view(1, 3)
    w, hw, hb = weights
    with torch.no_grad():
        v2, c2, m2 = ops.backproject_frames(WL["voxel_dim"], VS, origin, P[:, :2], feats_all[:2])
        vo, mo, co = O.encode_volume(WL["voxel_dim"], VS, origin, P[:, :2], [frame(0), frame(1)])
        lift_ok = bool(torch.equal(v2.cpu(), vo) and torch.equal(c2.cpu(), co) and torch.equal(m2.cpu(), mo))
        del v2, vo
        g = S.gen(99)
        idx = torch.randperm(xyz_h.shape[1], generator=g)[:n_check]
        vol_h = vol.cpu()
        valid_h = torch.ones((1, 1) + tuple(vol_h.shape[2:]), dtype=torch.bool)       # the lift zeroes unseen voxels itself
        pl_h = {k: planes[i].cpu() for i, k in enumerate(O.PLANES)}
        ref = O.gennerf_forward(xyz_h[:, idx], w, hw, hb, volume=vol_h, valid=valid_h, planes=pl_h, voxel_size=VS, padding=0.1,
                                num_freqs=MLP["num_freqs"], freq_factor=MLP["freq_factor"])
        err = (tsdf.reshape(1, -1, 1)[:, idx.to(dev)].cpu() - ref["tsdf"]).abs().max().item()
    return {"parity_checked": True, "lift_2_frames_full_grid_bit_exact": lift_ok, "tsdf_samples": n_check,
            "tsdf_max_abs_err_vs_fp32_oracle": err, "tsdf_bar": 1e-2, "ok": bool(lift_ok and err <= 1e-2)}


def run_native(args):
    import torch.distributed as dist
    from gennerf_b200 import ops, parallel
    from gennerf_b200.dropin import GenNerf
    from oracle.ref_shim import to_attr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # NCCL's banner / debug lines: not on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    T, H, W = WL["T"], WL["H"], WL["W"]
    nx, ny, nz = WL["voxel_dim"]
    V, Q = nx * ny * nz, WL["Q"]
    origin = torch.tensor([0, 0, 0]).view(1, 3)
    precision = args.precision

    P, pts_h, cpt_h, (w, hw, hb) = scene_small_parts()
    t0, t1 = parallel.shard_range(T, rank, world)
    q0, q1 = parallel.shard_range(Q, rank, world)
    n0, n1 = parallel.shard_range(pts_h.shape[1], rank, world)
    frames_h = [frame(t) for t in range(t0, t1)]                   # this rank's frames (its share of the CNN's output), NCHW
    xyz_h = queries(q0, q1)
    frames_d = [f.to(dev) for f in frames_h]
    xyz = xyz_h.to(dev)
    pts, cpt = pts_h[:, n0:n1].to(dev), cpt_h[:, n0:n1].to(dev)
    dw = ops.DecoderWeights(w, hw, hb, n_blocks=MLP["n_blocks"], d_geo=MLP["d_geo"], use_code=True, num_freqs=MLP["num_freqs"],
                            freq_factor=MLP["freq_factor"], device=dev)
    if precision != "fp32":
        dw.pack(precision)                                          # raises if the tcgen05 path is unavailable: no fallback
    fb = parallel.FrameBuffer(T, 1, C_FEAT, H, W, dev)
    # feature exchange: "p2p" = pulls over NVLink by the copy engines between symmetric-memory buffers, issued one scene ahead
    # (parallel.P2PFrameBuffer); "allgather" / "broadcast" = one NCCL collective inside the step.  "auto" takes p2p when every
    # rank could set it up.
    feat_mode, p2p = args.features, None
    if world > 1 and feat_mode in ("auto", "p2p"):
        ok = 1
        try:
            p2p = parallel.P2PFrameBuffer(T, 1, C_FEAT, H, W, dev)
        except Exception as e:                                      # noqa: BLE001  (no symmetric memory on this box / build)
            ok = 0
            print(f"[bench] rank {rank}: symmetric-memory exchange unavailable ({type(e).__name__}: {e}); NCCL all-gather instead",
                  file=sys.stderr)
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if feat_mode == "p2p":
                raise RuntimeError("--features p2p: symmetric memory could not be set up on every rank")
            p2p = None
        feat_mode = "p2p" if p2p is not None else "allgather"
    elif feat_mode == "auto":
        feat_mode = "allgather"
    ev = lambda: torch.cuda.Event(enable_timing=True)               # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Step:
        """One pass of the path; `marks` = events at the phase boundaries."""

        slot, ahead = 0, False                                      # p2p: the slot of this scene; is it already on its way?

        def __call__(self, marks=None, prefetch=False):
            """prefetch (p2p only): a further scene follows -- its frames are transposed and their exchange is started before
            this scene's lift, so that the copy engines move them while this scene's queries run."""
            def mark():
                if marks is not None:
                    e = ev()
                    e.record()
                    marks.append(e)
            mark()
            # ---- features: own frames NCHW -> NHWC into the flat buffer, all ranks' frames gathered over NVLink
            if p2p is not None:
                k = self.slot
                if not self.ahead:                                  # first scene of a stream: nothing was sent ahead
                    ops.nchw_to_nhwc(frames_d, out=p2p.own(k))
                    p2p.exchange(k)
                p2p.wait(k)
                self.ahead = bool(prefetch)
                if prefetch:
                    ops.nchw_to_nhwc(frames_d, out=p2p.own(k ^ 1))
                    p2p.exchange(k ^ 1)
                self.slot = k ^ 1
                frames = p2p.frames(k)
            else:
                if frames_d:
                    ops.nchw_to_nhwc(frames_d, out=fb.flat[t0:t1])
                if feat_mode == "broadcast" and world > 1:
                    # (alternative: rank 0 ran the CNN alone -- only its buffer is meaningful; one broadcast of 1.26 GB)
                    fb.broadcast(src=0)
                else:
                    fb.all_gather()
                frames = fb.frames
            mark()
            # ---- lift
            if args.lift == "slab" and world > 1:
                vol, cnt, valid = parallel.lift_sharded(ops, WL["voxel_dim"], VS, origin, P, frames, gather=True)
            else:
                vol, cnt, valid = ops.backproject_frames(WL["voxel_dim"], VS, origin, P, frames)
            mark()
            # ---- triplanes: own points -> partial sums -> all-reduce -> divide
            planes, pcnt = parallel.scatter_planes_sharded(ops, pts, cpt, R_PLANE, 0.1)
            pl = {k: planes[i] for i, k in enumerate(ops.PLANES)}
            mark()
            # ---- queries of this rank's range
            if precision == "fp32" or args.query == "unfused":
                feat = ops.sample_features(xyz, volume=vol, planes=pl, voxel_size=VS, origin=origin, padding=0.1)
                tsdf = ops.decode(dw, xyz, feat, precision)[1]
            else:
                # "image" = what the public op picks for this many queries: sampler kernel -> 16-bit operand image -> decoder
                tsdf = ops.query_fused(dw, xyz, volume=vol, planes=pl, voxel_size=VS, origin=origin, padding=0.1, want_feat=False,
                                       precision=precision, mode=args.query)[1]
            mark()
            self.out = (vol, cnt, planes, tsdf)
            return tsdf

    step = Step()
    if feat_mode == "broadcast" and world > 1 and rank == 0:        # rank 0 "ran the CNN": it holds every frame
        frames_all = [frame(t).to(dev) for t in range(T)]
        ops.nchw_to_nhwc(frames_all, out=fb.flat)
        del frames_all
    n_warm = max(args.warmup, 3)
    for i in range(n_warm):
        step(prefetch=p2p is not None and i + 1 < n_warm)
    barrier()

    clocks = ClockSampler(local)
    clocks.__enter__()                                              # sampled over every timed region below
    marks_all = []
    a, b = ev(), ev()
    barrier()
    t_wall0 = time.perf_counter()
    a.record()
    for i in range(args.steps):                                     # a stream of K scenes: every exchange is inside the region
        m = []
        step(m, prefetch=p2p is not None and i + 1 < args.steps)
        marks_all.append(m)
    b.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms_step = a.elapsed_time(b) / args.steps
    ph = torch.tensor([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in marks_all], dtype=torch.float64).mean(0)
    vol, cnt, planes, tsdf = step.out
    n_valid = int(cnt.sum().item())
    # the exchange timed alone (nothing to hide behind), max over ranks: the GB/s a GPU receives over NVLink
    ms_exchange_alone = None
    if world > 1:
        barrier()
        xa, xb = ev(), ev()
        xa.record()
        for i in range(3):
            if p2p is not None:
                p2p.exchange(i & 1)
                p2p.wait(i & 1)
            elif feat_mode == "broadcast":
                fb.broadcast(src=0)
            else:
                fb.all_gather()
        xb.record()
        barrier()
        tx = torch.tensor([xa.elapsed_time(xb) / 3], device=dev, dtype=torch.float64)
        dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        ms_exchange_alone = float(tx.item())
    overflow = dw.overflowed() if precision == "fp16" else False

    # ---- end to end through the drop-in API with pinned host buffers -----------------------------------------------
    cfg = to_attr({
        "voxel_size": VS, "voxel_dim_train": list(WL["voxel_dim"]), "voxel_dim_val": list(WL["voxel_dim"]),
        "voxel_dim_test": list(WL["voxel_dim"]),
        "encoder": {"use_spatial": True, "spatial": {"num_layers": 0, "latent_size": C_FEAT}, "use_pointnet": True, "use_auxiliary": False,
                    "pointnet": {"num_sparse_points": PTS_PER_FRAME, "c_dim": C_PLANE, "dim": 3, "padding": 0.1, "hidden_dim": 32,
                                 "scatter_type": "max", "plane_type": ["xz", "xy", "yz"], "plane_resolution": R_PLANE,
                                 "n_blocks": 5, "unet": False, "unet_kwargs": None, "sample_mode": "bilinear"},
                    "plane_merger": {"strategy": "average", "alpha": 0.1}},
        "mlp": {"d_out_sem": MLP["d_out"] - MLP["d_geo"], "d_out_geo": MLP["d_geo"], "n_blocks": MLP["n_blocks"],
                "d_hidden": MLP["d_hidden"], "combine_layer": 1000, "combine_type": "average", "beta": 0.0, "use_spade": False,
                "use_layer_norm": False, "alpha": 1.0},
        "use_code": True, "code": {"num_freqs": MLP["num_freqs"], "freq_factor": MLP["freq_factor"], "include_input": True}})
    torch.manual_seed(7)
    model = GenNerf(cfg, precision=precision).eval()
    model.mlp.load_state_dict(w)
    model.head_geo.load_state_dict({"fc.weight": hw, "fc.bias": hb})
    model = model.to(dev)
    if world > 1:
        model.shard_scene(p2p=feat_mode == "p2p")
    img_pin = torch.stack(frames_h, dim=1).pin_memory() if frames_h else torch.empty(1, 0, C_FEAT, H, W).pin_memory()
    xyz_pin = xyz_h.pin_memory()
    pts_pin = pts_h.pin_memory()
    tsdf_pin = torch.empty((1, q1 - q0, 1), dtype=torch.float32).pin_memory()
    # (the 32 camera matrices stay on the host: the lift takes them as kernel parameters, and a CUDA tensor would cost a
    # device-to-host copy, i.e. a stream sync, in the middle of every step)

    # Double-buffered, as a user streaming scenes through the drop-in would write it: scene i+1's inputs are uploaded on a
    # copy stream while scene i's kernels run, scene i's TSDF is downloaded on a third stream.  In the steady state the timed
    # region holds, per step, one full H2D of a scene's inputs, one pass of the hot path and one D2H of its result.
    cur, up, down = torch.cuda.current_stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def upload():
        with torch.cuda.stream(up):
            t = (img_pin.to(dev, non_blocking=True), xyz_pin.to(dev, non_blocking=True), pts_pin.to(dev, non_blocking=True))
            e = torch.cuda.Event()
            e.record(up)
        return t, e

    def e2e_step(pre):
        """pre = [scene i, scene i+1]: two scenes are on their way to the device.  Scene i+1's frames (uploaded during the previous
        step) are handed to the model before scene i is encoded: with the p2p exchange (N > 1) they cross NVLink under scene i's
        kernels; scene i+2's upload is issued here."""
        ((img, xd, sp), e), ((img_n, _, _), e_n) = pre
        cur.wait_event(e)
        cur.wait_event(e_n)
        nxt = upload()                                  # a further scene's host-to-device copies overlap this scene's kernels
        for t in (img, xd, sp, img_n):
            t.record_stream(cur)
        model.initialize_volume()
        with torch.no_grad():
            if world > 1:
                model.queue_next_frames(img_n)
            model.encode(P, img, None, "val", sparse_xyz=sp)
            out = model(xd)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(down):
            down.wait_event(done)
            tsdf_pin.copy_(out["tsdf"], non_blocking=True)
        out["tsdf"].record_stream(down)
        return [pre[1], nxt]

    pre = [upload(), upload()]
    for _ in range(2):
        pre = e2e_step(pre)
    cur.wait_stream(up)
    cur.wait_stream(down)
    barrier()
    a2, b2 = ev(), ev()
    a2.record()
    for _ in range(args.steps):
        pre = e2e_step(pre)
    cur.wait_stream(up)                                 # the K-th upload and download issued inside the region are inside it
    cur.wait_stream(down)
    b2.record()
    barrier()
    ms_e2e = a2.elapsed_time(b2) / args.steps
    clocks.__exit__(None, None, None)

    tt = torch.tensor([ms_step, ms_e2e] + ph.tolist(), device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step, ms_e2e, ms_feat, ms_lift, ms_planes, ms_query = tt.tolist()

    if rank == 0:
        pk = peaks()
        d_feat = C_FEAT + C_PLANE
        fl = flops_per_query(d_feat, D_CODE, MLP["d_hidden"], MLP["n_blocks"], MLP["d_out"], MLP["d_geo"]) * (q1 - q0)
        tf = fl / (ms_query * 1e-3) / 1e12
        lb = lift_bytes(T, C_FEAT, H, W, V, n_valid)
        lift_gbs = lb / (ms_lift * 1e-3) / 1e9
        feat_bytes = T * C_FEAT * H * W * 4
        qmode = "unfused" if precision == "fp32" else args.query
        n_chunks = -(-(q1 - q0) // ops.IMAGE_CHUNK)
        # own kernels per step: transpose, lift, scatter, finalize + the query phase (image: per chunk of 4 Mi queries the
        # brick sort [count, reduce, scan, scatter], the binned sampler and the decoder; fused: 1; unfused: sort + sampler + decoder)
        q_launches = {"image": 6 * n_chunks, "fused": 1, "unfused": 6}[qmode]
        q_desc = {"image": f"{precision} brick-binned sampler writing the decoder's 16-bit operand image + tcgen05 MLP decoder, "
                           f"{n_chunks} chunk(s) of <= {ops.IMAGE_CHUNK} queries",
                  "fused": f"{precision} tcgen05 fused sampler+MLP (one kernel)",
                  "unfused": f"{precision} sampler (fp32 features) + decoder kernels"}[qmode]
        line = {
            "metric": "tsdf_query_points_per_s", "value": Q / (ms_step * 1e-3), "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": {"fp16": "f16", "fp32": "f32"}[precision],
            "data": "synthetic", "config": config_dict(),
            "parallelism": {"gpus": world, "frames_per_rank": t1 - t0, "queries_per_rank": q1 - q0,
                            "features": {"p2p": "p2p pulls over NVLink by the copy engines (symmetric memory), issued one scene ahead: they run under "
                                                "the previous scene's query kernels; the first scene of the timed stream waits for its own",
                                         "allgather": "NCCL all-gather inside the step", "broadcast": "NCCL broadcast inside the step"}[feat_mode]
                            if world > 1 else "local", "lift": args.lift if world > 1 else "single GPU",
                            "planes": "points sharded, NCCL all-reduce of sums + counts", "queries": "contiguous ranges, no collective"},
            "decoder": q_desc,
            "phases_ms": {"features_transpose_and_gather": ms_feat, "lift": ms_lift, "planes_scatter_allreduce": ms_planes,
                          "query_range": ms_query, "step": ms_step},
            "collectives": {"feature_gather": {"bytes_total": feat_bytes, "bytes_received_per_gpu": feat_bytes * (world - 1) // world,
                                               "how": feat_mode if world > 1 else "local",
                                               "ms_in_step_incl_local_transpose": ms_feat,
                                               "ms_alone": ms_exchange_alone,
                                               "gbs_received_per_gpu": (feat_bytes * (world - 1) / world) / (ms_exchange_alone * 1e-3) / 1e9
                                               if world > 1 else None,
                                               "nvlink_reference_gbs": 770.0},
                            "plane_allreduce_bytes": 3 * R_PLANE * R_PLANE * (C_PLANE + 1) * 4},
            "roofline": {"kernel": "decoder_tc_kernel; timed = the whole query phase of this rank's range (" + q_desc + ")",
                         "bound": "tensor", "achieved": tf, "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": tf / pk["bf16_burst"],
                         "frac_of_sustained": tf / pk["bf16_sustained"], "traffic": None,
                         "peak_source": pk["source"] + ": bf16 cuBLAS burst (fp16 and bf16 share the tensor-core rate)",
                         "flops_per_launch": fl, "ms_per_launch": ms_query},
            "backprojection": {"metric": "voxel_frames_per_s", "value": V * T / (ms_lift * 1e-3), "ms": ms_lift,
                               "includes": "lift kernel on the gathered channels-last frames (every rank lifts the whole grid)",
                               "roofline": {"bound": "hbm", "achieved": lift_gbs, "peak": pk["hbm"], "unit": "GB/s",
                                            "frac": lift_gbs / pk["hbm"], "algorithmic_bytes": lb},
                               "valid_voxel_frames": n_valid},
            "timing": "K steps back to back between barriers, CUDA events, max over ranks; phases from events inside the steps",
            "e2e": {"value": Q / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": (img_pin.numel() + xyz_pin.numel() + pts_pin.numel()) * 4, "d2h_bytes_per_step": (q1 - q0) * 4,
                    "bytes_are": "per rank (every rank uploads its frames, its query range and the sparse points, downloads its TSDF range)",
                    "how": "drop-in GenNerf.shard_scene / encode(projection, image, sparse_xyz=) / forward(xyz) from pinned host "
                           "buffers: H2D, NCHW->NHWC, feature exchange (" + (feat_mode if world > 1 else "none at N = 1") + "), lift, PointNet + "
                           "scatter, query, D2H of the TSDF; scenes are streamed (uploads run two scenes ahead on a copy stream, the TSDF "
                           "downloads on a third stream; with the p2p exchange the next scene's frames cross NVLink under this scene's "
                           "kernels, model.queue_next_frames): per step the timed region holds one full H2D, one exchange, one pass and one D2H"},
            "gpu_launches": args.steps * (4 + q_launches + (1 if args.lift == "slab" and world > 1 else 0)),
            "fp16_saturated": overflow,
            "clocks": clocks.summary(),
            "wall_s_timed_region": t_wall,
        }
        if not args.no_parity:
            try:
                line["parity"] = parity_check(ops, dev, P, p2p.frames(0) if p2p is not None else fb.frames, vol, planes, tsdf, xyz_h, (w, hw, hb))
                line["parity_checked"] = line["parity"]["ok"]
            except Exception as e:                                  # noqa: BLE001
                line["parity"] = {"parity_checked": False, "error": repr(e)}
                line["parity_checked"] = False
        if world == 1 and not args.no_side:
            del step.out, vol, planes, tsdf
            try:
                line["hbm_kernels"] = hbm_side_kernels(ops, dev, pk)
            except Exception as e:                                  # a side measurement must not take the bench line down
                line["hbm_kernels"] = {"error": repr(e)}
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline(dev, P, pts_h, cpt_h, (w, hw, hb))
            except Exception as e:                                  # noqa: BLE001
                line["gpu_eager_baseline"] = {"error": repr(e)}
        if not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            t, detail = cpu_estimate(1, 10000, (P, pts_h, cpt_h, (w, hw, hb)))
            line["cpu_baseline"] = {"value": Q / t, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "lift of 1 of 32 frames into the full grid (scaled x32), the full triplane scatter, "
                                              "10 000 of 16 Mi queries as one reference chunk (scaled)", "detail": detail}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner, warnings)
    is sent to stderr; emit() writes the line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"],
                    help="decoder operands: fp16 = tcgen05 tensor cores (fp32 accumulate; the 1e-2 TSDF mode), fp32 = CUDA cores "
                         "(1e-5 mode)")
    ap.add_argument("--features", default="auto", choices=["auto", "p2p", "allgather", "broadcast"],
                    help="N > 1: every rank owns T/N frames.  p2p = every rank pulls the others' frames over NVLink with the copy "
                         "engines (symmetric memory), one scene ahead; allgather = one NCCL all-gather inside the step; broadcast = "
                         "rank 0 owns all frames and broadcasts; auto (default) = p2p when symmetric memory is available, else allgather")
    ap.add_argument("--lift", default="replicated", choices=["replicated", "slab"],
                    help="N > 1: every rank lifts the whole grid (default), or x-slabs + all-gather of the volume")
    ap.add_argument("--query", choices=["image", "fused", "unfused"], default="image",
                    help="query phase: image = sampler -> 16-bit operand image -> decoder (default, what ops.query_fused picks); "
                         "fused = one kernel; unfused = fp32 features between sampler and decoder")
    ap.add_argument("--unfused", action="store_true", help="same as --query unfused")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-side", action="store_true", help="skip the side measurements (HBM kernels, eager-GPU baseline)")
    ap.add_argument("--no-parity", action="store_true", help="skip the bench-time parity check")
    args = ap.parse_args()
    if args.unfused:
        args.query = "unfused"
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
