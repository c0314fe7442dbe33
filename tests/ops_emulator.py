"""TEST INFRASTRUCTURE ONLY -- a CPU stand-in for ``pde_policylearning_b200.ops`` (the ctypes layer over libb2no.so).

The product has no CPU path: every op of ``ops.py`` launches a CUDA kernel.  The HOST logic above it -- the hand-derived
backward passes of ``functional.py`` (which op is called with which operands, in which order, what is accumulated where)
and the module mirrors of ``modules.py`` -- is ordinary Python and can be checked without a GPU if the ops themselves
are replaced by their mathematical definition.  This module provides exactly that, built on the float64 oracle
(oracle/closed_form.py) and on the contract written in include/b2no.h; ``install()`` monkeypatches it
into the package for the duration of a test.  It is never imported by the product and nothing here is timed.
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as F

from oracle import closed_form as cf

RD, CD = torch.float64, torch.complex128


class _Plan:
    def __init__(self, geom):
        g = geom.resolved()
        self.geom = g
        self.cf = cf.SpecGeom(nin=g.nin, half=g.half, norm=g.norm, nfft=g.nfft, nout=g.nout)
        self.kept = self.cf.kept()
        self.modes = math.prod(self.kept)
        self.s_f, self.s_i = self.cf.scales()
        self.device = torch.device("cpu")
        self.layout = int(getattr(g, "layout", 0))

    def layout_supported(self, batch, channels):
        return self.layout == 0 or self.geom.ndim == 2      # exercise the mode-major host logic on the CPU too

    def spec_shape(self, batch, channels):
        return (self.kept + (batch, channels)) if self.layout else ((batch, channels) + self.kept)

    def spec_bc(self, spec):
        return (int(spec.shape[-2]), int(spec.shape[-1])) if self.layout else (int(spec.shape[0]), int(spec.shape[1]))

    def to_std(self, spec):
        """-> (batch, channel, *kept)"""
        if not self.layout:
            return spec
        d = len(self.kept)
        return spec.permute(d, d + 1, *range(d))

    def from_std(self, spec):
        if not self.layout:
            return spec
        d = len(self.kept)
        return spec.permute(*range(2, 2 + d), 0, 1).contiguous()


_plans = {}


def get_plan(geom, device=None):
    key = geom.resolved()
    if key not in _plans:
        _plans[key] = _Plan(geom)
    return _plans[key]


def _require_cuda(*tensors):
    return None


def _cplx(w):
    return w.to(CD) if w.is_complex() else torch.view_as_complex(w.detach().to(RD).contiguous())


def _act(z, act):
    if act in (None, "none", "identity"):
        return z
    return {"gelu": F.gelu, "relu": F.relu, "sigmoid": torch.sigmoid, "selu": F.selu, "tanh": torch.tanh}[act](z)


def _act_grad(z, act):
    if act in (None, "none", "identity"):
        return torch.ones_like(z)
    z = z.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        (g,) = torch.autograd.grad(_act(z, act).sum(), z)
    return g


def dft_forward(plan, which, x, out=None):
    spec = cf.dft_trunc(plan.cf, x, plan.s_f) if which == 0 else cf.dft_trunc(plan.cf, x, plan.s_i, use_inv_conj=True)
    spec = plan.from_std(spec.to(torch.complex64))
    if out is not None:
        out.copy_(spec)
        return out
    return spec


def make_epilogue(**kw):
    return dict(kw)


def _epilogue(z, epi, channels, grid):
    """z: float64 (B, Co, *grid) spectral term (or zeros); the contract of b2no_epilogue (include/b2no.h)."""
    epi = epi or {}
    nd = len(grid)
    B = z.shape[0]

    def pw(w, x, transposed):
        w = w.to(RD)
        w = w.t() if transposed else w
        return torch.einsum("oi,bi...->bo...", w, x.to(RD))

    if epi.get("bias") is not None:
        z = z + epi["bias"].to(RD).reshape((1, -1) + (1,) * nd)
    if epi.get("pw_w") is not None:
        z = z + pw(epi["pw_w"], epi["pw_x"], epi.get("pw_transposed", False))
    if epi.get("pw2_w") is not None:
        z = z + pw(epi["pw2_w"], epi["pw2_x"], epi.get("pw2_transposed", False))
    if epi.get("add") is not None:
        z = z + epi["add"].to(RD).reshape(z.shape)
    if epi.get("preact") is not None:
        epi["preact"].copy_(z.to(torch.float32).reshape(epi["preact"].shape))
    y = _act(z, epi.get("act"))
    if epi.get("mul") is not None:
        y = y * epi["mul"].to(RD).reshape(z.shape)
    if epi.get("dact_z") is not None:
        y = y * _act_grad(epi["dact_z"].to(RD).reshape(z.shape), epi.get("dact"))
    if epi.get("gate_z") is not None:
        y = y + (1.0 - epi["gate_z"].to(RD).reshape(z.shape)) * epi["gate_h"].to(RD).reshape(z.shape)
    return y.to(torch.float32)


def dft_inverse(plan, which, spec, epi=None, out=None):
    spec = plan.to_std(spec)
    z = cf.idft_trunc(plan.cf, spec, plan.s_i) if which == 0 else cf.idft_trunc(plan.cf, spec, plan.s_f, use_fwd_conj=True)
    grid = plan.geom.nout if which == 0 else plan.geom.nin
    y = _epilogue(z, epi, spec.shape[1], tuple(grid))
    if out is not None:
        out.copy_(y)
        return out
    return y


def pointwise(batch, channels, grid, device, epi):
    z = torch.zeros((batch, channels) + tuple(grid), dtype=RD)
    return _epilogue(z, epi, channels, tuple(grid))


def mix(plan, mode, spec, corners, ci, co, out=None, accumulate=False):
    W = cf.gather_weight(plan.cf, [_cplx(c) for c in corners])
    s = plan.to_std(spec).to(CD)
    if mode == 0:
        r = torch.einsum("bi...,io...->bo...", s, W)
    else:
        r = torch.einsum("bo...,io...->bi...", s, W.conj())
    r = plan.from_std(r.to(torch.complex64))
    if out is not None:
        if accumulate:
            out.add_(r)
        else:
            out.copy_(r)
        return out
    return r


def mix_dw(plan, xh, gyh, like, needs_zero, out=None, accumulate=False):
    dW_eff = torch.einsum("bi...,bo...->io...", plan.to_std(xh).to(CD).conj(), plan.to_std(gyh).to(CD))
    shapes = [tuple(t.shape if t.is_complex() else t.shape[:-1]) for t in like]
    dW = cf.scatter_weight_grad(plan.cf, dW_eff, shapes)
    res = []
    for i, (t, d) in enumerate(zip(like, dW)):
        d = d.to(torch.complex64) if t.is_complex() else torch.view_as_real(d).to(torch.float32)
        if out is not None:
            if accumulate:
                out[i].add_(d)
            else:
                out[i].copy_(d)
            res.append(out[i])
        else:
            res.append(d.contiguous())
    return res


def act_bwd(gy, z, act):
    return (gy.to(RD) * _act_grad(z.to(RD), act)).to(torch.float32)


def pw_wgrad(g, x, need_bias, db_out=None):
    dw = torch.einsum("bo...,bi...->oi", g.to(RD), x.to(RD)).to(torch.float32)
    db = None
    if need_bias:
        db = g.to(RD).sum(dim=[0] + list(range(2, g.dim()))).to(torch.float32)
        if db_out is not None:
            db_out.copy_(db)
            db = db_out
    return dw, db


def mlp_head_bwd_supported(ci, hidden, pixels):
    return False


def mlp_head_bwd_fused_supported(ci, hidden, pixels):
    return ci <= 32 and hidden <= 256 and pixels % 128 == 0


def mlp_head_bwd_fused(x, w1, b1, w2, g, act="gelu", dact_z=None, dact=None):
    """Definition of b2no_mlp_head_bwd_fused (include/b2no.h): (gx, dW1, db1, dw2, db2) of out = w2 . act(W1 x + b1)."""
    nd = x.dim() - 2
    xd, w1d, w2d = (t.detach().to(RD).requires_grad_(True) for t in (x, w1, w2))
    b1d = (torch.zeros(w1.shape[0], dtype=RD) if b1 is None else b1.detach().to(RD)).requires_grad_(True)
    with torch.enable_grad():
        z = torch.einsum("ji,bi...->bj...", w1d, xd) + b1d.reshape((1, -1) + (1,) * nd)
        y = torch.einsum("j,bj...->b...", w2d, _act(z, act)).unsqueeze(1)
        gx, dw1, db1, dw2 = torch.autograd.grad(y, [xd, w1d, b1d, w2d], g.to(RD))
    if dact_z is not None:
        zz = dact_z.detach().to(RD).requires_grad_(True)
        with torch.enable_grad():
            (da,) = torch.autograd.grad(_act(zz, dact).sum(), zz)
        gx = gx * da
    return tuple(t.to(torch.float32) for t in (gx, dw1, db1, dw2, g.to(RD).sum().reshape(1)))


def mlp_head_fwd(x, w1, b1, w2, b2, act="gelu"):
    z = torch.einsum("ji,bi...->bj...", w1.to(RD), x.to(RD))
    nd = x.dim() - 2
    if b1 is not None:
        z = z + (b1.to(RD).reshape((1, -1) + (1,) * nd) if b1.dim() == 1 else b1.to(RD).reshape(b1.shape + (1,) * nd))
    y = torch.einsum("j,bj...->b...", w2.to(RD), _act(z, act)).unsqueeze(1)
    if b2 is not None:
        y = y + b2.to(RD).reshape(())
    return y.to(torch.float32)


def rno_gate_fwd(z, z2, hh, h):
    return (1.0 - z) * h + z2 * hh


def rno_gate_bwd(g, z, z2, hh, h):
    return [-g * h, g * hh, g * z2, g * (1.0 - z)]


def rno_cell_bwd(g, h, zz2, ah, g_zz2, g_ah):
    C = h.shape[1]
    g, h, zz2, ah = (t.to(RD) for t in (g, h, zz2, ah))
    z, z2 = zz2[:, :C], zz2[:, C:]
    hh = F.selu(ah)
    g_zz2[:, :C].copy_((-g * h * z * (1 - z)).to(torch.float32))
    g_zz2[:, C:].copy_((g * hh * z2 * (1 - z2)).to(torch.float32))
    g_ah.copy_((g * z2 * _act_grad(ah, "selu")).to(torch.float32))
    return (g * (1 - z)).to(torch.float32)


def rno_reset_bwd(g_rh, h, ar, g_ar, g_h):
    r = torch.sigmoid(ar.to(RD))
    g_ar.copy_((g_rh.to(RD) * h.to(RD) * r * (1 - r)).to(torch.float32))
    g_h.add_((g_rh.to(RD) * r).to(torch.float32))


def rel_l2_sums(x, y):
    d = (x.to(RD) - y.to(RD)).square().sum(dim=1)
    n = y.to(RD).square().sum(dim=1)
    return torch.stack((d, n), dim=1).to(torch.float32)


def rel_l2_finish(sums, size_average):
    s = sums.to(RD)
    r = s[:, 0].sqrt() / s[:, 1].sqrt()
    loss = r.mean() if size_average else r.sum()
    scale = 1.0 / sums.shape[0] if size_average else 1.0
    coef = scale / (s[:, 0].sqrt() * s[:, 1].sqrt())
    return loss.to(torch.float32), coef.to(torch.float32)


def rel_l2_bwd_g(x, y, coef, g):
    return (g.to(RD) * coef.to(RD)[:, None] * (x.to(RD) - y.to(RD))).to(torch.float32)


def pino_residual_supported(n):
    return 8 <= n <= 64 and n % 4 == 0


def _pino_losses(w, u0, forcing2d, nu, t_interval):
    from pde_policylearning_b200 import pino_loss
    B, N, _, T = w.shape
    du = pino_loss.fdm_ns_vorticity(w, nu, t_interval)                     # the GEMM composition, float64 here
    f = forcing2d.reshape(1, N, N, 1).expand(B, N, N, T - 2)
    rel = lambda a, b: (torch.linalg.vector_norm((a - b).reshape(B, -1), dim=1) / torch.linalg.vector_norm(b.reshape(B, -1), dim=1)).mean()
    return torch.stack((rel(w[..., 0], u0), rel(du, f)))


def pino_residual_fwd(w, u0, forcing2d, nu, t_interval):
    with torch.no_grad():
        loss2 = _pino_losses(w.to(RD), u0.to(RD), forcing2d.to(RD), nu.to(RD), t_interval).to(torch.float32)
    return loss2, None, None, None


def pino_residual_bwd(w, u0, forcing2d, nu, t_interval, du_p, fields, coef, gup):
    w64 = w.to(RD).detach().requires_grad_(True)
    with torch.enable_grad():
        loss2 = _pino_losses(w64, u0.to(RD), forcing2d.to(RD), nu.to(RD), t_interval)
        (dw,) = torch.autograd.grad((loss2 * gup.to(RD)).sum(), w64)
    return dw.to(torch.float32)


_NAMES = ["pino_residual_supported", "pino_residual_fwd", "pino_residual_bwd", "get_plan", "_require_cuda", "dft_forward", "make_epilogue", "dft_inverse", "pointwise", "mix", "mix_dw",
          "act_bwd", "pw_wgrad", "mlp_head_bwd_supported", "mlp_head_bwd_fused_supported", "mlp_head_bwd_fused", "mlp_head_fwd", "rno_gate_fwd", "rno_gate_bwd", "rno_cell_bwd",
          "rno_reset_bwd", "rel_l2_sums", "rel_l2_finish", "rel_l2_bwd_g"]


@contextlib.contextmanager
def installed():
    """Replace the CUDA ops by their definitions for the duration of the block (host-logic tests on the CPU)."""
    import pde_policylearning_b200.functional as Fn
    import pde_policylearning_b200.ops as ops
    extra = {}
    saved = {n: getattr(ops, n) for n in _NAMES if hasattr(ops, n)}
    saved_fn_plan = Fn.get_plan
    saved_supported = getattr(Fn, "rno_layer_supported", None)
    g = globals()
    try:
        for n in _NAMES:
            setattr(ops, n, g[n])
        Fn.get_plan = get_plan
        if saved_supported is not None:
            Fn.rno_layer_supported = lambda x, C, H, W: H == W and (C * H * W) % 4 == 0
        for k, v in extra.items():
            setattr(ops, k, v)
        yield
    finally:
        for n, v in saved.items():
            setattr(ops, n, v)
        Fn.get_plan = saved_fn_plan
        if saved_supported is not None:
            Fn.rno_layer_supported = saved_supported
