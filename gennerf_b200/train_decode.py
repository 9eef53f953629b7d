"""Training-time decoder on the tensor cores (nn.Linear forward / backward of ResnetFC under
autograd, reference resnetfc.py:134-189 + heads3d.py:36-50).

Forward = the tcgen05 decoder kernel (fp16 operands, fp32 accumulation) with `gnb_decode_tc_save`: it also stores the fp16
activations every layer consumed.  Backward uses exactly those: the ReLU masks are `activation > 0`, the weight gradients are
`grad^T @ activation`.  The gradient GEMMs run in fp32 storage through torch.matmul (TF32 tensor cores under
torch.set_float32_matmul_precision("high"), the reference's own training setting, src/utils/utils.py:48) -- no loss scaling
is needed because no gradient is ever stored in 16 bits.

The gradients are those of the network the kernel evaluated (fp16-rounded operands); against fp32 autograd of the fp32
network they differ by the fp16 rounding of the activations (tests/test_gpu_train_decode.py states the bar).
"""
import os

import torch
from torch.autograd.function import once_differentiable

from . import ops
from .torch_ops import mlp_keys


# backward of the decoder: True = ONE library call (gnb_decode_train_bwd: own kernels + cuBLAS SGEMMs behind the C ABI);
# False = the same chain spelled out here with torch.matmul between the own kernels (what the tests compare it with)
NATIVE_BACKWARD = os.environ.get("GNB_TRAIN_BWD", "native") != "python"

_pending_status = []          # (pinned host copy of a forward's status word, event after the copy), not read yet


def _watch_status(status):
    """Queues an asynchronous copy of a forward's device status word into pinned host memory; check_saturation reads it once
    the copy's event has completed -- the training step itself never waits for the device."""
    host = torch.empty(1, dtype=torch.int32, pin_memory=True)
    host.copy_(status, non_blocking=True)
    status.zero_()                                # the weights view (and its status word) lives on across steps
    ev = torch.cuda.Event()
    ev.record()
    _pending_status.append((host, ev))


def check_saturation(wait=False):
    """Raises if an fp16 operand of an earlier training forward saturated at +-65504 (that step's gradients are those of a
    clipped network).  Called at the start of every decode_train_tc: only status words whose copy has already arrived are
    looked at (no device sync inside a step, so the report can come a step or two late); `wait=True` -- call it yourself
    after the last step -- waits for all of them."""
    hit = False
    for item in list(_pending_status):
        host, ev = item
        if wait:
            ev.synchronize()
        elif not ev.query():
            continue
        _pending_status.remove(item)
        hit |= bool(host.item())
    if hit:
        raise FloatingPointError("gennerf_b200: an fp16 operand of a training forward saturated at +-65504; "
                                 "use train_precision='fp32' for this model")


_views = {}                   # (device, d_code, parameter addresses) -> DecoderWeights: pointers only, rebuilt when a tensor moves


def _weights_view(code, head_w, head_b, n_blocks, d_geo, params):
    """The pointer view of the live fp32 parameters (ops.DecoderWeights) is the same object step after step as long as the
    optimiser updates the parameters in place; only the 16-bit operand image is re-packed per forward."""
    ok = all(p.dtype == torch.float32 and p.is_contiguous() and p.device == code.device for p in (head_w, head_b, *params))
    key = (code.device, code.shape[1], n_blocks, d_geo, head_w.data_ptr(), head_b.data_ptr(), *(p.data_ptr() for p in params)) if ok else None
    dw = _views.get(key) if ok else None
    if dw is None:
        sd = dict(zip(mlp_keys(n_blocks), params))
        dw = ops.DecoderWeights(sd, head_w, head_b, n_blocks=n_blocks, d_geo=d_geo, use_code=2, num_freqs=0, freq_factor=0.0,
                                include_input=False, d_code=code.shape[1], device=code.device, alpha_on_device=True)
        if ok:
            if len(_views) >= 8:
                _views.clear()
            _views[key] = dw
    return dw


class _DecodeTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, code, feat, head_w, head_b, n_blocks, d_geo, *params):
        dw = _weights_view(code, head_w, head_b, n_blocks, d_geo, params)
        dw.pack("fp16")                              # the parameters changed since the last step: new operand image
        out, tsdf, acts = ops.decode_save(dw, code, feat, "fp16")
        ctx.n_blocks, ctx.d_geo = n_blocks, d_geo
        ctx.dw = dw
        ctx.set_materialize_grads(False)             # an output the loss does not use arrives as None, not as a tensor of zeros
        _watch_status(dw.status)                     # read at a later call, when its copy has arrived: no stall of the launch queue
        ctx.save_for_backward(code, feat, out, tsdf, acts, head_w, *params)
        return out, tsdf

    @staticmethod
    def backward(ctx, g_out, g_tsdf):
        # create_graph=True (the eikonal / gradient losses, reference utils.py:636-649) runs this with grad mode ON.
        # once_differentiable alone only raises when the INCOMING gradients carry history; the dependence of this backward
        # on the saved activations would be dropped without a word.
        if torch.is_grad_enabled():
            raise RuntimeError("gennerf_b200: the fp16 training decoder (train_precision='fp16') is once-differentiable; "
                               "use train_precision='fp32' with loss.use_eikonal / loss.use_gradient (create_graph=True)")
        return _DecodeTC._backward_once(ctx, g_out, g_tsdf)

    @staticmethod
    @once_differentiable
    def _backward_once(ctx, g_out, g_tsdf):
        code, feat, out, tsdf, acts, head_w, *params = ctx.saved_tensors
        nb, d_geo = ctx.n_blocks, ctx.d_geo
        keys = mlp_keys(nb)
        if NATIVE_BACKWARD:
            need = ctx.needs_input_grad
            grads, d_hw, d_hb, g_code, g_feat = ops.decode_train_bwd(ctx.dw, code, feat, out, tsdf, acts, g_out, g_tsdf,
                                                                     need_code=need[0], need_feat=need[1])
            plist = [grads[k].reshape(p.shape) if need[6 + j] else None for j, (k, p) in enumerate(zip(keys, params))]
            return (g_code, g_feat, d_hw.reshape(head_w.shape) if (d_hw is not None and need[2]) else None,
                    d_hb if (d_hb is not None and need[3]) else None, None, None, *plist)
        P = dict(zip(keys, params))
        alpha = P["alpha"]
        grads = {k: None for k in P}
        n, H = acts.shape[1], acts.shape[2]
        dc = code.shape[1]
        # every link below = one library GEMM (grad @ W, fp32 storage / TF32) + ONE pass of gnb_mlp_grad_link (ReLU mask from
        # the saved 16-bit activation, skip-connection add, bias column sum, fp32 copy of the activation for the wgrad GEMM)
        G, d_hw, d_hb, d_lob = ops.mlp_grad_head(g_out, g_tsdf, out, tsdf, head_w, d_geo)
        colsum = torch.zeros((2 * nb + 1, H), device=code.device, dtype=torch.float32)
        S, a32 = ops.mlp_grad_link(G @ P["lin_out.weight"], acts[2 * nb], colsum=colsum[2 * nb], want_act32=True)
        grads["lin_out.weight"] = G.t() @ a32
        grads["lin_out.bias"] = d_lob
        # S = gradient w.r.t. the residual stream; the one entering block i's lin_z sits in column slab i of GX, so that the
        # five K = d_code GEMMs of lin_z (weights and codes) are ONE GEMM each over the concatenation
        GX = torch.empty((n, max(nb, 1) * H), device=code.device, dtype=torch.float32)
        for i in reversed(range(nb)):
            grads[f"blocks.{i}.fc_1.bias"] = colsum[2 * i + 2]
            gn, h32 = ops.mlp_grad_link(S @ P[f"blocks.{i}.fc_1.weight"], acts[2 * i + 1], colsum=colsum[2 * i + 1], want_act32=True)
            grads[f"blocks.{i}.fc_1.weight"] = S.t() @ h32
            grads[f"blocks.{i}.fc_0.bias"] = colsum[2 * i + 1]
            Sn = GX[:, i * H:(i + 1) * H]
            _, a32 = ops.mlp_grad_link(gn @ P[f"blocks.{i}.fc_0.weight"], acts[2 * i], S, out=Sn, colsum=colsum[2 * i], want_act32=True)
            grads[f"blocks.{i}.fc_0.weight"] = gn.t() @ a32
            S = Sn
        g_code = None
        if nb:
            pad = (-dc) % 4                                              # K = 15 would put cuBLAS on its unaligned kernels
            code_p = torch.nn.functional.pad(code, (0, pad))
            wz = torch.nn.functional.pad(torch.cat([P[f"lin_z.{i}.weight"] for i in range(nb)], 0), (0, pad))   # (nb*H, dc+pad)
            gz = GX @ wz                                                  # (n, dc+pad) = sum_i S_i @ Wz_i
            g_code = (gz[:, :dc] * alpha) if ctx.needs_input_grad[0] else None
            d_wz = ((GX.t() @ code_p)[:, :dc] * alpha).reshape(nb, H, dc) # contiguous (H, dc) gradients per block
            gsum = colsum[0:2 * nb:2]                                     # (nb, H): column sums of S_i
            for i in range(nb):
                grads[f"lin_z.{i}.weight"] = d_wz[i]
            d_zb = gsum * alpha
            for i in range(nb):
                grads[f"lin_z.{i}.bias"] = d_zb[i]
            bz = torch.stack([P[f"lin_z.{i}.bias"] for i in range(nb)], 0)
            grads["alpha"] = ((gz * code_p).sum() + (gsum * bz).sum()).reshape(alpha.shape)
        else:
            grads["alpha"] = torch.zeros_like(alpha)
        grads["lin_in.weight"] = S.t() @ feat
        grads["lin_in.bias"] = colsum[0]
        need = ctx.needs_input_grad
        g_feat = (S @ P["lin_in.weight"]) if need[1] else None
        if g_code is None and need[0]:
            g_code = torch.zeros_like(code)
        plist = [grads[k] if need[6 + j] else None for j, k in enumerate(keys)]
        has_t = g_tsdf is not None
        return (g_code if need[0] else None, g_feat, d_hw.reshape(head_w.shape) if (has_t and need[2]) else None,
                d_hb if (has_t and need[3]) else None, None, None, *plist)


def decode_train_tc(mlp, head, z, feat):
    """z (..., d_code) positional codes (with autograd history back to xyz), feat (..., C_lat) -> out (..., d_out),
    tsdf (..., 1) through the tcgen05 kernel, differentiable once w.r.t. z, feat and every ResnetFC / head parameter."""
    check_saturation()
    lead = z.shape[:-1]
    sd = dict(mlp.named_parameters())
    params = [sd[k] for k in mlp_keys(mlp.n_blocks)]
    out, tsdf = _DecodeTC.apply(z.reshape(-1, z.shape[-1]).float(), feat.reshape(-1, feat.shape[-1]).float(), head.fc.weight,
                                head.fc.bias, mlp.n_blocks, head.fc.weight.shape[1], *params)
    return out.reshape(*lead, -1), tsdf.reshape(*lead, 1)
