"""PINO PDE-residual loss of the channel-flow observer without torch.fft / cuFFT.

Reference: ``FDM_NS_vorticity`` (libs/envs/diff_control_env.py:5-41) and ``Channelflow_PINO_loss`` (:44-60):
residual ``w_t + u . grad w - nu lap w`` with spectral derivatives over (x, y) and a central difference in t,
compared to the forcing by the relative-L2 loss of libs/pino_utils/losses.py:182-194.

The reference takes a full ``fft2`` of every time slice and five ``irfft2`` of the half spectrum.  Here the same
sums are written as contractions with the DFT matrices (the form every transform of this package has):

    A[x, ky]   = sum_y w[x, y] e^{-2 pi i ky y / N}              ky = 0 .. N/2      (real -> complex, half spectrum)
    W[kx, ky]  = sum_x e^{-2 pi i kx x / N} A[x, ky]             kx = 0 .. N-1
    G[x, ky]   = 1/N sum_kx e^{+2 pi i kx x / N} m(kx, ky) W[kx, ky]
    r[x, y]    = 1/N sum_ky c(ky) Re( G[x, ky] e^{+2 pi i ky y / N} )   c(0) = c(N/2) = 1, else 2   (C2R: Im of the ky = 0 and
                                                                                                      Nyquist columns dropped)

as plain real matrix products (library GEMMs) over the whole (B, T) batch, the five multipliers m stacked so that the two
inverse contractions run once.  Quirks kept: the signed wavenumber of index N/2 is -N/2 on both axes (:15-19), the (0, 0)
entry of the Laplacian is set to 1 for EVERY use, including ``-lap * w_h`` (:21-29).  Autograd differentiates the
composition; the relative-L2 reductions use the package's loss kernels on CUDA tensors.

Two implementations of the same sums:
  * the FUSED kernels of csrc/pino_loss.cu (``PinoResidualLossFn``: one CTA per (sample, time slice), spectral derivatives,
    products, central difference, warp-shuffle loss partials and the hand-derived backward; square grids 8 <= N <= 64,
    N % 4 == 0 -- the BASELINE cfg4 grid is 64 x 64) -- what ``channelflow_pino_loss`` runs on CUDA tensors;
  * the composition of library GEMMs with the DFT matrices below (``fdm_ns_vorticity``), kept as the differentiable
    utility for larger grids and as the second opinion the GPU tests compare the kernels with."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import functional as Fn
from . import ops

_tables: Dict[tuple, tuple] = {}


def _dft_tables(n: int, device, dtype) -> Tuple[torch.Tensor, ...]:
    """cos / sin tables for the full axis (n x n) and the half-spectrum axis ((n/2 + 1) x n), built in float64."""
    key = (n, str(device), dtype)
    t = _tables.get(key)
    if t is None:
        k = torch.arange(n, dtype=torch.float64)
        ang = 2.0 * math.pi * torch.outer(k, k) / n               # [k, x]
        cx, sx = torch.cos(ang), torch.sin(ang)
        h = n // 2 + 1
        t = tuple(m.to(device=device, dtype=dtype).contiguous() for m in (cx, sx, cx[:h], sx[:h]))
        _tables[key] = t
    return t


def fdm_ns_vorticity(w: torch.Tensor, v: torch.Tensor, t_interval: float = 1.0) -> torch.Tensor:
    """Residual of the 2-D vorticity equation without the forcing term (diff_control_env.py:5-41).

    w: (B, N, N, T) real; v: (B,) viscosities (1 / Re).  Returns (B, N, N, T - 2)."""
    B, nx, ny, nt = w.shape
    if nx != ny or nx % 2:
        raise ValueError("the PINO residual needs a square grid of even size (diff_control_env.py:13-19)")
    if w.is_cuda and w.dtype == torch.float32 and ops.pino_residual_supported(nx) and not (
            torch.is_grad_enabled() and (w.requires_grad or v.requires_grad)):
        # no gradient wanted: the fused kernel's residual planes (B, T, N, N) -> the reference's (B, N, N, T - 2)
        wc = w.contiguous()
        zero = torch.zeros((B, nx, ny), dtype=torch.float32, device=w.device)
        _, _, du_p, _ = ops.pino_residual_fwd(wc, zero, zero[0].contiguous(), v.reshape(B).float().contiguous(), t_interval)
        return du_p[:, 1:nt - 1].permute(0, 2, 3, 1)
    n, kmax, h = nx, nx // 2, nx // 2 + 1
    dt_, dev = w.dtype, w.device
    cx, sx, cy, sy = _dft_tables(n, dev, dt_)                      # cx, sx: [kx, x]; cy, sy: [ky <= N/2, y]
    # forward, y axis (real -> complex half spectrum), then x axis; layout (B, T, x | kx, y | ky)
    wt_ = w.permute(0, 3, 1, 2)                                    # (B, T, x, y) view
    ar = torch.matmul(wt_, cy.t())                                 # sum_y w cos
    ai = -torch.matmul(wt_, sy.t())                                # -sum_y w sin
    wr = torch.matmul(cx, ar) + torch.matmul(sx, ai)               # (cos - i sin)(ar + i ai)
    wi = torch.matmul(cx, ai) - torch.matmul(sx, ar)
    # signed wavenumbers: index N/2 carries -N/2 on both axes (:15-19)
    k1 = torch.cat((torch.arange(0, kmax), torch.arange(-kmax, 0))).to(device=dev, dtype=dt_)
    kx = k1.reshape(n, 1)
    ky = k1[:h].reshape(1, h)
    lap = kx ** 2 + ky ** 2
    lap = lap.clone()
    lap[0, 0] = 1.0
    fr, fi = wr / lap, wi / lap                                    # stream function f_h = w_h / lap
    # the five spectra, multiplied by i k (.) -> (re, im) = (-k im, k re)
    spec_r = torch.stack((-ky * fi, kx * fi, -kx * wi, -ky * wi, -lap * wr), dim=0)    # ux, uy, wx, wy, wlap
    spec_i = torch.stack((ky * fr, -kx * fr, kx * wr, ky * wr, -lap * wi), dim=0)
    # inverse, x axis (complex), scale 1/N
    gr = (torch.matmul(cx.t(), spec_r) - torch.matmul(sx.t(), spec_i)) / n            # (cos + i sin)(re + i im), [x, kx] = table^T
    gi = (torch.matmul(cx.t(), spec_i) + torch.matmul(sx.t(), spec_r)) / n
    # inverse, y axis (C2R): c(ky) Re(G e^{+i}) = c (gr cos - gi sin), scale 1/N
    c = torch.full((h, 1), 2.0, device=dev, dtype=dt_)
    c[0, 0] = 1.0
    c[kmax, 0] = 1.0
    fields = (torch.matmul(gr, c * cy) - torch.matmul(gi, c * sy)) / n                # (5, B, T, x, y)
    ux, uy, wx, wy, wlap = fields.unbind(0)
    nu = v.reshape(-1, 1, 1, 1).to(dt_)
    adv = (ux * wx + uy * wy - nu * wlap).permute(0, 2, 3, 1)                          # back to (B, x, y, T)
    dt = t_interval / (nt - 1)
    w_t = (w[:, :, :, 2:] - w[:, :, :, :-2]) / (2 * dt)
    return w_t + adv[..., 1:-1]


def get_forcing(S: int, device=None, dtype=torch.float32) -> torch.Tensor:
    """libs/pino_utils/losses.py:288-291: -4 cos(4 y) on the periodic grid, shape (1, S, S, 1)."""
    # the grid is rounded to fp32 BEFORE the cosine, as the reference does (np.linspace in float64 -> torch.float):
    # bit-identical forcing, including its ~1e-7 values where the exact cosine vanishes
    x2 = (torch.arange(S, dtype=torch.float64) * (2.0 * math.pi / S)).to(torch.float32)
    f = -4 * torch.cos(4 * x2)
    return f.reshape(1, 1, S, 1).repeat(1, S, 1, 1).to(device=device, dtype=dtype)


def _rel(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """LpLoss(size_average=True).rel, p = 2 (libs/pino_utils/losses.py:182-194)."""
    if x.is_cuda:
        return Fn.rel_l2_loss(x.reshape(x.shape[0], -1), y.reshape(y.shape[0], -1), True)
    raise RuntimeError("pde_policylearning_b200: the PINO loss reductions run only on CUDA (no CPU fallback)")


class PinoResidualLossFn(torch.autograd.Function):
    """[loss_ic, loss_f] = Channelflow_PINO_loss(w, u0, forcing, nu, t_interval) on the fused kernels (csrc/pino_loss.cu)."""

    @staticmethod
    def forward(ctx, w, u0, forcing2d, nu, t_interval):
        c = lambda t: t.float().contiguous()
        w, u0, forcing2d, nu = c(w), c(u0), c(forcing2d), c(nu)
        loss2, coef, du_p, fields = ops.pino_residual_fwd(w, u0, forcing2d, nu, t_interval)
        ctx.t_interval = float(t_interval)
        ctx.save_for_backward(w, u0, forcing2d, nu, du_p, fields, coef)
        return loss2

    @staticmethod
    def backward(ctx, g):
        w, u0, forcing2d, nu, du_p, fields, coef = ctx.saved_tensors
        dw = ops.pino_residual_bwd(w, u0, forcing2d, nu, ctx.t_interval, du_p, fields, coef, g.float().contiguous())
        return dw, None, None, None, None


def channelflow_pino_loss(model_output: torch.Tensor, u0: torch.Tensor, forcing: torch.Tensor, v: torch.Tensor,
                          t_interval: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss_ic, loss_f) of diff_control_env.py:44-60: initial-condition and PDE-residual relative-L2 losses."""
    B, nx, ny, nt = model_output.shape[:4]
    out = model_output.reshape(B, nx, ny, nt)
    if out.is_cuda and nx == ny and ops.pino_residual_supported(nx) and forcing.numel() == nx * ny:
        loss2 = PinoResidualLossFn.apply(out, u0.reshape(B, nx, ny), forcing.reshape(nx, ny), v.reshape(B), t_interval)
        return loss2[0], loss2[1]
    loss_ic = _rel(out[:, :, :, 0], u0)
    du = fdm_ns_vorticity(out, v, t_interval)
    f = forcing.expand(B, nx, ny, nt - 2)
    return loss_ic, _rel(du, f)


def pino_training_loss(model, a_in: torch.Tensor, re: torch.Tensor, u: torch.Tensor, forcing: torch.Tensor,
                       xy_weight: float = 5.0, f_weight: float = 1.0, ic_weight: float = 1.0,
                       t_interval: float = 1.0) -> torch.Tensor:
    """The loss of one PINO training iteration, train_pino.py:87-107:

        loss = xy_weight * LpLoss(out, u) + f_weight * loss_f + ic_weight * loss_ic

    The reference runs the model TWICE per iteration when both terms are on (train_pino.py:88 and :98, quirk Q6): same
    weights, same input, no dropout -- the two outputs are the same tensor, so it is computed once here and both losses
    (and both gradient contributions) use it."""
    B, S1, S2, T = a_in.shape[:4]
    out = model(a_in, re).reshape(B, S1, S2, T)
    loss = None
    if xy_weight > 0:
        loss = xy_weight * _rel(out, u)
    if f_weight != 0.0:
        u0 = a_in[:, :, :, 0, -1]
        loss_ic, loss_f = channelflow_pino_loss(out, u0, forcing, 1.0 / re, t_interval)
        extra = f_weight * loss_f + ic_weight * loss_ic
        loss = extra if loss is None else loss + extra
    return loss
