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
import torch
from torch.autograd.function import once_differentiable

from . import ops
from .torch_ops import mlp_keys


_pending_status = []          # status words of earlier forwards, not read yet


def check_saturation():
    """Raises if an fp16 operand of an earlier training forward saturated at +-65504 (that step's gradients are those of a
    clipped network).  Called at the start of every decode_train_tc, i.e. one step late but without a device sync inside
    the step; call it yourself after the last step."""
    hit = False
    while _pending_status:
        hit |= bool(_pending_status.pop().item())
    if hit:
        raise FloatingPointError("gennerf_b200: an fp16 operand of a training forward saturated at +-65504; "
                                 "use train_precision='fp32' for this model")


class _DecodeTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, code, feat, head_w, head_b, n_blocks, d_geo, *params):
        sd = dict(zip(mlp_keys(n_blocks), params))
        dw = ops.DecoderWeights(sd, head_w, head_b, n_blocks=n_blocks, d_geo=d_geo, use_code=2, num_freqs=0, freq_factor=0.0,
                                include_input=False, d_code=code.shape[1], device=code.device)
        out, tsdf, acts = ops.decode_save(dw, code, feat, "fp16")
        ctx.n_blocks, ctx.d_geo = n_blocks, d_geo
        _pending_status.append(dw.status)            # read at the NEXT call (by then the step is over: no stall of the launch queue)
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
        P = dict(zip(mlp_keys(nb), params))
        alpha = P["alpha"]
        grads = {k: None for k in P}
        f32 = torch.float32
        G = torch.zeros_like(out) if g_out is None else g_out.to(f32).clone()
        d_hw = d_hb = None
        if g_tsdf is not None:
            s = g_tsdf.to(f32) * (1.0 - tsdf * tsdf)                     # (n,1): through tanh
            d_hw = s.t() @ out[:, :d_geo]                                # (1,d_geo)
            d_hb = s.sum(0)
            G[:, :d_geo] += s @ head_w.reshape(1, -1)
        relu_bwd = torch.ops.aten.threshold_backward          # grad * (activation > 0) in one kernel
        a_f = acts[2 * nb].to(f32)
        grads["lin_out.weight"] = G.t() @ a_f
        grads["lin_out.bias"] = G.sum(0)
        gx = relu_bwd(G @ P["lin_out.weight"], a_f, 0.0)                 # grad wrt x_nb
        g_code = torch.zeros_like(code)
        d_alpha = torch.zeros((), device=code.device, dtype=f32)
        for i in reversed(range(nb)):
            h, a = acts[2 * i + 1].to(f32), acts[2 * i].to(f32)
            grads[f"blocks.{i}.fc_1.weight"] = gx.t() @ h
            grads[f"blocks.{i}.fc_1.bias"] = gx.sum(0)
            gn = relu_bwd(gx @ P[f"blocks.{i}.fc_1.weight"], h, 0.0)
            grads[f"blocks.{i}.fc_0.weight"] = gn.t() @ a
            grads[f"blocks.{i}.fc_0.bias"] = gn.sum(0)
            gx = gx + relu_bwd(gn @ P[f"blocks.{i}.fc_0.weight"], a, 0.0)  # grad wrt u_i = x_i + alpha * lin_z_i(code)
            gsum = gx.sum(0)
            Wz, bz = P[f"lin_z.{i}.weight"], P[f"lin_z.{i}.bias"]
            grads[f"lin_z.{i}.weight"] = alpha * (gx.t() @ code)
            grads[f"lin_z.{i}.bias"] = alpha * gsum
            gz = gx @ Wz                                                  # (n, d_code)
            g_code.add_(gz, alpha=1.0)
            d_alpha = d_alpha + (gz * code).sum() + (gsum * bz).sum()
        g_code = g_code * alpha
        grads["lin_in.weight"] = gx.t() @ feat
        grads["lin_in.bias"] = gx.sum(0)
        g_feat = gx @ P["lin_in.weight"]
        grads["alpha"] = d_alpha.reshape(alpha.shape)
        need = ctx.needs_input_grad
        plist = [grads[k] if need[6 + j] else None for j, k in enumerate(mlp_keys(nb))]
        return (g_code if need[0] else None, g_feat if need[1] else None,
                d_hw.reshape(head_w.shape) if (d_hw is not None and need[2]) else None,
                d_hb if (d_hb is not None and need[3]) else None, None, None, *plist)


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
