"""torch.autograd.Function wrappers over the training-path kernels (csrc/train.cu, include/dedf.h "training path").

The reference trains through torch autograd over e3nn / torch_scatter ops (trainer.py:308-346).  Here autograd only
orchestrates: every forward and every backward below is a hand-written CUDA kernel; torch is used for memory movement
(slice / cat / reshape / transpose) and for accumulating gradients of tensors that are used more than once.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import _lib as L
from . import ops
from ._lib import ptr, stream

Irr = Tuple[int, int, int]


def _dim(irr: Irr) -> int:
    return irr[0] + 3 * irr[1] + 5 * irr[2]


def _split(w_flat: torch.Tensor, irr_in: Irr, irr_out: Irr):
    """Flat LinearRS weight -> per-l (m_in, m_out) views (None where a block is absent)."""
    out, off = [], 0
    for a, b in zip(irr_in, irr_out):
        if a and b:
            out.append(w_flat[off:off + a * b].view(a, b))
            off += a * b
        else:
            out.append(None)
    assert off == w_flat.numel(), (off, w_flat.numel(), irr_in, irr_out)
    return out


class LinearFn(Function):
    """y_l = W_l^T x_l per l (+ bias on the scalars): LinearRS / FCTP with 1x0e / nn.Linear (irreps (K,0,0) -> (N,0,0))."""

    @staticmethod
    def forward(ctx, x, w_flat, bias, irr_in: Irr, irr_out: Irr):
        x = x.contiguous()
        Ws = [w.contiguous() if w is not None else None for w in _split(w_flat.detach(), irr_in, irr_out)]
        y = ops.node_linear(x, irr_in, irr_out, Ws, bias.detach().contiguous() if bias is not None else None)
        ctx.save_for_backward(x, w_flat)
        ctx.irr = (tuple(irr_in), tuple(irr_out))
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w_flat = ctx.saved_tensors
        irr_in, irr_out = ctx.irr
        gy = gy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            WsT = [w.t().contiguous() if w is not None else None for w in _split(w_flat.detach(), irr_in, irr_out)]
            dx = ops.node_linear(gy, irr_out, irr_in, WsT, None)
        if not (ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])):
            return dx, None, None, None, None          # frozen parameters (EbmScoreModelHead.forward in inference mode)
        dW = torch.zeros_like(w_flat)
        db = torch.zeros(irr_out[0], dtype=torch.float32, device=x.device) if ctx.has_bias else None
        dWs = _split(dW, irr_in, irr_out)
        ops._call("dedf_lin_wgrad", ptr(x), ptr(gy), x.shape[0], L.int_array(irr_in), L.int_array(irr_out),
                  ptr(dWs[0]), ptr(dWs[1]), ptr(dWs[2]), ptr(db), stream())
        return dx, dW, db, None, None


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, irr: Irr, eps: float):
        x = x.contiguous()
        y = torch.empty_like(x)
        ops._call("dedf_ln_fwd", ptr(x), x.shape[0], L.int_array(irr), ptr(w.detach().contiguous()),
                  ptr(b.detach().contiguous()) if b is not None else None, eps, ptr(y), stream())
        ctx.save_for_backward(x, w)
        ctx.irr, ctx.eps, ctx.has_b = tuple(irr), eps, b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w)
        db = torch.zeros(max(ctx.irr[0], 1), dtype=torch.float32, device=x.device)
        ops._call("dedf_ln_bwd", ptr(x), ptr(g), x.shape[0], L.int_array(ctx.irr), ptr(w.detach().contiguous()), ctx.eps,
                  ptr(dx), ptr(dw), ptr(db), stream())
        return dx, dw, (db[:ctx.irr[0]] if ctx.has_b else None), None, None


class GateFn(Function):
    """Gate: ``irr_pre`` = pre-gate irreps (m0 = scalars + gates)."""

    @staticmethod
    def forward(ctx, pre, irr_pre: Irr):
        pre = pre.contiguous()
        fy = _dim(irr_pre) - irr_pre[1] - irr_pre[2]
        y = torch.empty(pre.shape[0], fy, dtype=torch.float32, device=pre.device)
        ops._call("dedf_gate_fwd", ptr(pre), pre.shape[0], L.int_array(irr_pre), ptr(y), stream())
        ctx.save_for_backward(pre)
        ctx.irr = tuple(irr_pre)
        return y

    @staticmethod
    def backward(ctx, g):
        (pre,) = ctx.saved_tensors
        d = torch.empty_like(pre)
        ops._call("dedf_gate_bwd", ptr(pre), ptr(g.contiguous()), pre.shape[0], L.int_array(ctx.irr), ptr(d), stream())
        return d, None


class ActFn(Function):
    """Elementwise activation: mode 0 SiLU (nn.SiLU, no normalisation constant), mode 1 sigmoid."""

    @staticmethod
    def forward(ctx, x, mode: int):
        x = x.contiguous()
        y = torch.empty_like(x)
        ops._call("dedf_act_fwd", ptr(x), x.numel(), mode, ptr(y), stream())
        ctx.save_for_backward(x)
        ctx.mode = mode
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        d = torch.empty_like(x)
        ops._call("dedf_act_bwd", ptr(x), ptr(g.contiguous()), x.numel(), ctx.mode, ptr(d), stream())
        return d, None


class SiluFn:
    @staticmethod
    def apply(x):
        return ActFn.apply(x, 0)


class SigmoidFn:
    @staticmethod
    def apply(x):
        return ActFn.apply(x, 1)


class DtpFn(Function):
    """Depthwise tensor product with the harmonics: (E,F) x (E,9) x weights -> (E, 49 G).  ``w``: (E, 15 G) or (15 G,)."""

    @staticmethod
    def forward(ctx, x, sh, w, mul1: int):
        x, sh, w = x.contiguous(), sh.contiguous(), w.contiguous()
        E = x.shape[0]
        shared = w.dim() == 1
        out = torch.empty(E, 49 * mul1, dtype=torch.float32, device=x.device)
        ops._call("dedf_dtp_fwd", mul1, ptr(x), ptr(sh), ptr(w.detach()), 0 if shared else w.shape[1], E, ptr(out), stream())
        ctx.save_for_backward(x, sh, w)
        ctx.mul1, ctx.shared = mul1, shared
        return out

    @staticmethod
    def backward(ctx, g):
        x, sh, w = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w) if ctx.shared else torch.empty_like(w)
        ops._call("dedf_dtp_bwd", ctx.mul1, ptr(x), ptr(sh), ptr(w.detach()), 0 if ctx.shared else w.shape[1], ptr(g), x.shape[0],
                  ptr(dx), ptr(dw), stream())
        dsh = None
        if ctx.needs_input_grad[1]:        # position gradients (EbmScoreModelHead.forward): the harmonics depend on the query coordinates
            dsh = torch.empty_like(sh)
            ops._call("dedf_dtp_bwd_sh", ctx.mul1, ptr(x), ptr(w.detach()), 0 if ctx.shared else w.shape[1], ptr(g), x.shape[0],
                      ptr(dsh), stream())
        return dx, dsh, dw, None


def dtp_generic(x: torch.Tensor, sh: torch.Tensor, w: torch.Tensor, irr_in: Irr, paths, irr_out: Irr) -> torch.Tensor:
    """Depthwise tensor product for irreps outside the fused kernels' family (table-driven, irreps.dtp_paths).  FORWARD ONLY: the
    training path (a backward through it) exists for the 2G:G:G/2 family of the shipped configs."""
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or sh.requires_grad):
        raise NotImplementedError("gradients through the table-driven depthwise tensor product (irreps outside 2G x0e + G x1e + G/2 x2e) "
                                  "are not built")
    x, sh, w = x.contiguous(), sh.contiguous(), w.contiguous()
    E = x.shape[0]
    shared = w.dim() == 1
    out = torch.empty(E, _dim(irr_out), dtype=torch.float32, device=x.device)
    flat = [int(v) for p in paths for v in p]
    ops._call("dedf_dtp_generic_fwd", ptr(x), L.int_array(irr_in), ptr(sh), ptr(w), 0 if shared else w.shape[1], len(paths),
              L.int_array(flat), L.int_array(irr_out), E, ptr(out), stream())
    return out


def dtp(sep, x: torch.Tensor, sh: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """The depthwise tensor product of a SeparableFCTP: the family-specialised kernels (with backward) or the table-driven one."""
    if sep.fused_family:
        return DtpFn.apply(x, sh, w, sep.irreps_node.m[1])
    return dtp_generic(x, sh, w, sep.irreps_node.m, sep.paths, sep.irreps_dtp_out.m)


class GatherFn(Function):
    """y = x[idx] (idx int32 or int64); backward scatters with atomics."""

    @staticmethod
    def forward(ctx, x, idx):
        x = x.contiguous()
        y = torch.empty(idx.shape[0], x.shape[1], dtype=torch.float32, device=x.device)
        if idx.dtype == torch.int32:
            ops._call("dedf_gather_rows_i32", ptr(x), ptr(idx, torch.int32), idx.shape[0], x.shape[1], ptr(y), stream())
        else:
            ops._call("dedf_gather_rows", ptr(x), ptr(idx, torch.long), idx.shape[0], x.shape[1], ptr(y), stream())
        ctx.save_for_backward(idx)
        ctx.n = x.shape[0]
        return y

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        out = torch.zeros(ctx.n, g.shape[1], dtype=torch.float32, device=g.device)
        ops._call("dedf_scatter_add_rows", ptr(g), idx.data_ptr(), 1 if idx.dtype == torch.long else 0, idx.shape[0], g.shape[1],
                  ptr(out), stream())
        return out, None


class AddScaleFn(Function):
    """(a + b) * s  (residuals, the UNet's (a+b)/sqrt(3) skips)."""

    @staticmethod
    def forward(ctx, a, b, s: float):
        ctx.s = s
        return ops.add_scale(a.contiguous(), b.contiguous(), s)

    @staticmethod
    def backward(ctx, g):
        if ctx.s == 1.0:
            return g, g, None
        g = g.contiguous()
        gs = ops.add_scale(g, g, 0.5 * ctx.s)          # (g + g) * s/2 = g * s
        return gs, gs, None


class AlphaFn(Function):
    @staticmethod
    def forward(ctx, pre, alpha_dot, edge_logit):
        pre = pre.contiguous()
        E, ma = pre.shape
        logits = torch.empty(E, 4, dtype=torch.float32, device=pre.device)
        ad = alpha_dot.detach().reshape(-1).contiguous()
        ops._call("dedf_alpha_fwd", ptr(pre), E, ma, ptr(ad), ptr(edge_logit), ptr(logits), stream())
        ctx.save_for_backward(pre, alpha_dot)
        return logits

    @staticmethod
    def backward(ctx, g):
        pre, alpha_dot = ctx.saved_tensors
        E, ma = pre.shape
        dpre = torch.empty_like(pre)
        dad = torch.zeros(ma, dtype=torch.float32, device=pre.device)
        ops._call("dedf_alpha_bwd", ptr(pre), E, ma, ptr(alpha_dot.detach().reshape(-1).contiguous()), ptr(g.contiguous()),
                  ptr(dpre), ptr(dad), stream())
        dlogit = None
        if ctx.needs_input_grad[2]:        # logits[e, h] = ... + edge_logit[e]
            g = g.contiguous()
            dlogit = torch.empty(E, dtype=torch.float32, device=pre.device)
            ones = torch.ones_like(g)
            ops._call("dedf_rowdot", ptr(g), ptr(ones), E, g.shape[1], ptr(dlogit), stream())
        return dpre, dad.view_as(alpha_dot), dlogit


class SoftmaxReduceFn(Function):
    @staticmethod
    def forward(ctx, logits, val, g: ops.Csr, irr: Irr):
        logits, val = logits.contiguous(), val.contiguous()
        out = ops.segment_softmax_reduce(g, logits, val, irr)
        ctx.save_for_backward(logits, val)
        ctx.g, ctx.irr = g, tuple(irr)
        return out

    @staticmethod
    def backward(ctx, gout):
        logits, val = ctx.saved_tensors
        g = ctx.g
        dl, dv = torch.zeros_like(logits), torch.zeros_like(val)
        ops._call("dedf_softmax_reduce_bwd", ptr(g.row_ptr, torch.int32), g.n_dst, g.n_seg, ptr(logits), ptr(val),
                  ptr(gout.contiguous()), ctx.irr[0], ctx.irr[1], ctx.irr[2], ptr(dl), ptr(dv), stream())
        return dl, dv, None, None


class RbfFn(Function):
    """Gaussian radial basis of edge lengths with learnable (mean, std_logit, weight_logit), each of shape (1, K)."""

    @staticmethod
    def forward(ctx, length, mean, std_logit, weight_logit, offset: float, inv_span: float, mode: int):
        length = length.contiguous()
        K = mean.numel()
        out = torch.empty(length.shape[0], K, dtype=torch.float32, device=length.device)
        p = [t.detach().reshape(-1).contiguous() for t in (mean, std_logit, weight_logit)]
        ops._call("dedf_rbf_fwd", ptr(length), length.shape[0], K, ptr(p[0]), ptr(p[1]), ptr(p[2]), offset, inv_span, mode, ptr(out), stream())
        ctx.save_for_backward(length, mean, std_logit, weight_logit)
        ctx.cfg = (offset, inv_span, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        length, mean, std_logit, weight_logit = ctx.saved_tensors
        K = mean.numel()
        p = [t.detach().reshape(-1).contiguous() for t in (mean, std_logit, weight_logit)]
        d = [torch.zeros(K, dtype=torch.float32, device=length.device) for _ in range(3)]
        ops._call("dedf_rbf_bwd", ptr(length), length.shape[0], K, ptr(p[0]), ptr(p[1]), ptr(p[2]), *ctx.cfg, ptr(g.contiguous()),
                  ptr(d[0]), ptr(d[1]), ptr(d[2]), stream())
        dlen = None
        if ctx.needs_input_grad[0]:
            dlen = torch.empty_like(length)
            ops._call("dedf_rbf_bwd_len", ptr(length), length.shape[0], K, ptr(p[0]), ptr(p[1]), ptr(p[2]), *ctx.cfg, ptr(g.contiguous()),
                      ptr(dlen), stream())
        return dlen, d[0].view_as(mean), d[1].view_as(std_logit), d[2].view_as(weight_logit), None, None, None


def sinusoid(x: torch.Tensor, freq: torch.Tensor, dim: int, scale: float) -> torch.Tensor:
    """SinusoidalPositionEmbeddings (no parameters, no gradient w.r.t. x on this path)."""
    x = x.detach().contiguous()
    out = torch.empty(x.shape[0], dim, dtype=torch.float32, device=x.device)
    ops._call("dedf_sinusoid", ptr(x), x.shape[0], dim, ptr(freq), scale, ptr(out), stream())
    return out


class SinusoidFn(Function):
    """SinusoidalPositionEmbeddings of a quantity that carries a gradient (the edge lengths of the all-pairs scale on the
    EbmScoreModelHead.forward path)."""

    @staticmethod
    def forward(ctx, x, freq, dim: int, scale: float):
        x = x.contiguous()
        out = torch.empty(x.shape[0], dim, dtype=torch.float32, device=x.device)
        ops._call("dedf_sinusoid", ptr(x), x.shape[0], dim, ptr(freq), scale, ptr(out), stream())
        ctx.save_for_backward(x, freq)
        ctx.cfg = (dim, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        x, freq = ctx.saved_tensors
        dx = torch.empty_like(x)
        ops._call("dedf_sinusoid_bwd", ptr(x), x.shape[0], ctx.cfg[0], ptr(freq), ctx.cfg[1], ptr(g.contiguous()), ptr(dx), stream())
        return dx, None, None, None


class EdgeGeomFn(Function):
    """graph_parser.py:146-224 with a gradient w.r.t. the DESTINATION (query) coordinates: x_dst -> (length (E), harmonics (E, 9),
    edge logit (E)).  The sources (the encoded scene) are constants on the EbmScoreModelHead.forward path."""

    @staticmethod
    def forward(ctx, x_dst, x_src, g: ops.Csr, radii, src_off, ns_cut):
        x_dst, x_src = x_dst.contiguous(), x_src.contiguous()
        length, sh, logit = ops.edge_geom(x_src, x_dst, g, radii=radii, src_off=src_off, ns_cut=ns_cut, want_logit=True)
        E = g.n_edges
        ctx.save_for_backward(x_dst, x_src)
        ctx.cfg = (g, list(radii), list(src_off), tuple(ns_cut))
        return length[:E].clone(), sh[:E].clone(), logit[:E].clone()

    @staticmethod
    def backward(ctx, g_len, g_sh, g_logit):
        x_dst, x_src = ctx.saved_tensors
        g, radii, src_off, ns_cut = ctx.cfg
        E = g.n_edges
        dx = torch.zeros_like(x_dst)
        z = lambda t, shape: (t.contiguous() if t is not None else torch.zeros(shape, dtype=torch.float32, device=x_dst.device))   # noqa: E731
        g_len, g_sh, g_logit = z(g_len, (E,)), z(g_sh, (E, 9)), z(g_logit, (E,))
        ops._call("dedf_edge_geom_bwd", ptr(x_src), ptr(x_dst), ptr(g.edge_src, torch.int32), ptr(g.edge_dst, torch.int32), E,
                  len(radii), L.int_array(src_off), L.float_array([-1.0 if r is None else float(r) for r in radii]),
                  float(ns_cut[0]), float(ns_cut[1]), ptr(g_len), ptr(g_sh), ptr(g_logit), ptr(dx), stream())
        return dx, None, None, None, None, None


class EbmEnergyFn(Function):
    """energy[t] = scale sum_q w_q |key_f[t, q] - query_f[t, q]|^2 (score_head_ebm.py:171-172) with gradients to both features."""

    @staticmethod
    def forward(ctx, key_f, query_f, qw, n_t: int, n_q: int, scale: float):
        key_f, query_f, qw = key_f.contiguous(), query_f.contiguous(), qw.contiguous()
        ctx.save_for_backward(key_f, query_f, qw)
        ctx.cfg = (n_t, n_q, scale)
        return ops.ebm_energy(key_f, query_f, qw, n_t, n_q, scale)

    @staticmethod
    def backward(ctx, g):
        key_f, query_f, qw = ctx.saved_tensors
        n_t, n_q, scale = ctx.cfg
        dk, dq = torch.empty_like(key_f), torch.empty_like(query_f)
        ops._call("dedf_ebm_energy_bwd", ptr(key_f), ptr(query_f), ptr(qw), ptr(g.contiguous()), n_t, n_q, key_f.shape[1], scale,
                  ptr(dk), ptr(dq), stream())
        return dk, dq, None, None, None, None


class ScoreTpFn(Function):
    @staticmethod
    def forward(ctx, a, b, w, irr: Irr):
        a, b = a.contiguous(), b.contiguous()
        m0, m1, m2 = irr
        out = torch.empty(a.shape[0], (m0 + m1 + m2) + 3 * (m0 + 3 * m1 + 2 * m2), dtype=torch.float32, device=a.device)
        ops._call("dedf_score_tp_fwd", ptr(a), ptr(b), ptr(w.detach().contiguous()), a.shape[0], L.int_array(irr), ptr(out), stream())
        ctx.save_for_backward(a, b, w)
        ctx.irr = tuple(irr)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, w = ctx.saved_tensors
        da, db, dw = torch.zeros_like(a), torch.zeros_like(b), torch.zeros_like(w)
        ops._call("dedf_score_tp_bwd", ptr(a), ptr(b), ptr(w.detach().contiguous()), a.shape[0], L.int_array(ctx.irr),
                  ptr(g.contiguous()), ptr(da), ptr(db), ptr(dw), stream())
        return da, db, dw, None


class QueryTransformFn(Function):
    """f' = D(q_t) f_q for every (pose, query point); gradient w.r.t. the query features only."""

    @staticmethod
    def forward(ctx, Ts, qx, qf, irr: Irr):
        Ts, qx, qf = Ts.contiguous(), qx.contiguous(), qf.contiguous()
        _, f = ops.query_transform(Ts, qx, qf.detach(), irr)
        ctx.save_for_backward(Ts)
        ctx.irr, ctx.shape = tuple(irr), qf.shape
        return f

    @staticmethod
    def backward(ctx, g):
        (Ts,) = ctx.saved_tensors
        n_q, F = ctx.shape
        dqf = torch.zeros(n_q, F, dtype=torch.float32, device=g.device)
        ops._call("dedf_query_transform_bwd", ptr(Ts), Ts.shape[0], n_q, L.int_array(ctx.irr), ptr(g.contiguous()), ptr(dqf), stream())
        return None, None, dqf, None


class AssembleFn(Function):
    """score_head.py:196-209: (gated lin / ang outputs, poses, query coords / weights) -> (ang (nT,3), lin (nT,3))."""

    @staticmethod
    def forward(ctx, ylin, yang, Ts, qx, qw, n_vec: int, lin_mult: float):
        ylin, yang, Ts, qx, qw = [t.contiguous() for t in (ylin, yang, Ts, qx, qw)]
        n_t, n_q = Ts.shape[0], qx.shape[0]
        ang = torch.empty(n_t, 3, dtype=torch.float32, device=Ts.device)
        lin = torch.empty(n_t, 3, dtype=torch.float32, device=Ts.device)
        ops._call("dedf_assemble_fwd", ptr(Ts), n_t, n_q, n_vec, ptr(ylin), ptr(yang), ptr(qx), ptr(qw.detach()), lin_mult,
                  ptr(ang), ptr(lin), stream())
        ctx.save_for_backward(ylin, yang, Ts, qx, qw)
        ctx.cfg = (n_vec, lin_mult)
        return ang, lin

    @staticmethod
    def backward(ctx, gang, glin):
        ylin, yang, Ts, qx, qw = ctx.saved_tensors
        n_vec, lin_mult = ctx.cfg
        dyl, dya = torch.empty_like(ylin), torch.empty_like(yang)
        dqw = torch.zeros_like(qw)
        ops._call("dedf_assemble_bwd", ptr(Ts), Ts.shape[0], qx.shape[0], n_vec, ptr(ylin), ptr(yang), ptr(qx), ptr(qw.detach()),
                  lin_mult, ptr(gang.contiguous()), ptr(glin.contiguous()), ptr(dyl), ptr(dya), ptr(dqw), stream())
        return dyl, dya, None, None, dqw, None, None


# ------------------------------------------------------------------------------------------------ train-mode dropout
def dropout_mask(shape, p: float, device) -> torch.Tensor:
    """Philox mask of 0 / 1/(1-p).  Every mask takes a fresh 62-bit Philox seed from torch's CPU generator (no device
    synchronisation), so torch.manual_seed() makes a training run reproducible."""
    seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    n = 1
    for v in shape:
        n *= int(v)
    out = torch.empty(*shape, dtype=torch.float32, device=device)
    ops._call("dedf_dropout_mask", seed, 0, n, float(p), ptr(out), stream())
    return out


class GroupScaleFn(Function):
    """x * mask broadcast per attention head (mode 0) or per irrep channel (mode 1); the mask is a constant."""

    @staticmethod
    def forward(ctx, x, mask, irr: Irr, mode: int):
        x, mask = x.contiguous(), mask.contiguous()
        y = torch.empty_like(x)
        ops._call("dedf_group_scale", ptr(x), ptr(mask), x.shape[0], L.int_array(irr), mode, ptr(y), stream())
        ctx.save_for_backward(mask)
        ctx.cfg = (tuple(irr), mode)
        return y

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        g = g.contiguous()
        d = torch.empty_like(g)
        ops._call("dedf_group_scale", ptr(g), ptr(mask), g.shape[0], L.int_array(ctx.cfg[0]), ctx.cfg[1], ptr(d), stream())
        return d, None, None, None


class RowScaleFn(Function):
    """y[r, :] = x[r, :] * s[r, 0] with gradients to both (source-point attention: s = w[edge_src], graph_attention.py:258-259)."""

    @staticmethod
    def forward(ctx, x, s, irr: Irr):
        x, s = x.contiguous(), s.contiguous()
        y = torch.empty_like(x)
        ops._call("dedf_group_scale", ptr(x), ptr(s), x.shape[0], L.int_array(irr), 2, ptr(y), stream())
        ctx.save_for_backward(x, s)
        ctx.irr = tuple(irr)
        return y

    @staticmethod
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(x)
        ops._call("dedf_group_scale", ptr(g), ptr(s), g.shape[0], L.int_array(ctx.irr), 2, ptr(dx), stream())
        ds = torch.empty_like(s)
        ops._call("dedf_rowdot", ptr(g), ptr(x), x.shape[0], x.shape[1], ptr(ds), stream())
        return dx, ds, None
