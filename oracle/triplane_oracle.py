"""CPU oracle for the tri-plane volume-rendering hot path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch numpy/float32 restatement of the algorithm the
reference implements with PyTorch ops.  It is *not* part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.
The product path (``g-nerf_b200``) never imports anything from ``oracle/``.

Parity pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against the reference *itself*, executed on CPU in the
build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/g_nerf``); the resulting known-answer tensors are committed
under ``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.

All paths below are relative to /root/reference/g_nerf/ ;
VR/ = training/volumetric_rendering/ .

Arithmetic contract (what "the oracle's result" means where association order
matters):
  * everything is IEEE float32 unless stated;
  * the importance-sampling CDF is built with float64 accumulation rounded to
    float32 per entry -- this is bit-identical to torch's CPU ``cumsum`` (probed)
    and, because every partial sum is exact in float64 for weights in the range
    the renderer produces, it is independent of summation order, so a parallel
    GPU scan reproduces it bit for bit;
  * the row sum that normalises the pdf is likewise float64-accumulated and then
    rounded once to float32 (torch's CPU ``sum`` uses a machine-dependent
    vectorised float32 order; it differs from this by at most one ulp).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# a1. RaySampler.forward                                  VR/ray_sampler.py:24-63
# --------------------------------------------------------------------------
def ray_sample(cam2world: np.ndarray, intrinsics: np.ndarray, resolution: int):
    """Pixel-centre rays from cam2world [N,4,4] and normalised intrinsics [N,3,3].

    Follows VR/ray_sampler.py:36-61: pixel centres (i+0.5)/res with x fastest
    (:43-44), lift with fx, fy, cx, cy, skew (:51-52), rotate by cam2world (:56),
    subtract the camera position, normalise (:58-59), origin repeated (:61).
    """
    c2w = np.asarray(cam2world, F32)
    K = np.asarray(intrinsics, F32)
    n = c2w.shape[0]
    res = int(resolution)
    fx, fy = K[:, 0, 0][:, None], K[:, 1, 1][:, None]
    cx, cy = K[:, 0, 2][:, None], K[:, 1, 2][:, None]
    sk = K[:, 0, 1][:, None]
    ticks = np.arange(res, dtype=F32) * F32(1.0 / res) + F32(0.5 / res)
    x_cam = np.tile(ticks, res)[None, :].repeat(n, 0)          # x fastest
    y_cam = np.repeat(ticks, res)[None, :].repeat(n, 0)
    x_lift = (x_cam - cx + cy * sk / fy - sk * y_cam / fy) / fx
    y_lift = (y_cam - cy) / fy
    ones = np.ones_like(x_lift)
    pts = np.stack([x_lift, y_lift, ones, ones], -1).astype(F32)   # [N,M,4]
    # torch.bmm accumulates k = 0..3 in float32
    world = np.zeros((n, res * res, 3), F32)
    for i in range(3):
        acc = np.zeros((n, res * res), F32)
        for k in range(4):
            acc = acc + c2w[:, i, k][:, None] * pts[:, :, k]
        world[:, :, i] = acc
    cam = c2w[:, :3, 3][:, None, :]
    d = world - cam
    nrm = np.sqrt((d * d).sum(-1, keepdims=True, dtype=F32)).astype(F32)
    d = d / np.maximum(nrm, F32(1e-12))                            # F.normalize eps
    o = np.broadcast_to(cam, d.shape).astype(F32).copy()
    return o, d.astype(F32)


# --------------------------------------------------------------------------
# a2/a3. generate_planes + project_onto_planes              VR/renderer.py:23-53
# --------------------------------------------------------------------------
# The three axis triples at VR/renderer.py:29-37 are permutation matrices; the
# reference multiplies coordinates by their inverses and keeps two columns
# (:51-53).  The result (probed against the reference) is a pure selection:
#   plane 0 <- (x, y), plane 1 <- (x, z), plane 2 <- (z, x);
# the first selected coordinate indexes the plane's W axis, the second its H.
PLANE_UV = ((0, 1), (0, 2), (2, 0))


def project_onto_planes(xyz: np.ndarray) -> np.ndarray:
    """[N,P,3] -> [N,3,P,2] plane coordinates (u -> width, v -> height)."""
    return np.stack([xyz[..., list(sel)] for sel in PLANE_UV], 1)


# --------------------------------------------------------------------------
# a4. sample_from_planes (bilinear, zeros padding)           VR/renderer.py:55-65
# --------------------------------------------------------------------------
def gather_planes(planes: np.ndarray, xyz: np.ndarray, box_warp: float) -> np.ndarray:
    """planes [N,3,C,H,W], xyz [N,P,3] -> features [N,3,P,C].

    Coordinates are scaled by 2/box_warp (VR/renderer.py:61) and looked up with
    grid_sample(bilinear, zeros, align_corners=False) (:64): pixel position
    ((g+1)*size-1)/2, four taps weighted nw, ne, sw, se, accumulated in that
    order; taps outside the plane contribute zero.
    """
    planes = np.asarray(planes, F32)
    n, n_pl, c, h, w = planes.shape
    g = project_onto_planes(np.asarray(xyz, F32) * F32(2.0 / box_warp))     # [N,3,P,2]
    out = np.zeros((n, n_pl, xyz.shape[1], c), F32)
    for i in range(n):
        for p in range(n_pl):
            gx, gy = g[i, p, :, 0], g[i, p, :, 1]
            ix = ((gx + F32(1)) * F32(w) - F32(1)) / F32(2)
            iy = ((gy + F32(1)) * F32(h) - F32(1)) / F32(2)
            x0f, y0f = np.floor(ix), np.floor(iy)
            x1f, y1f = x0f + F32(1), y0f + F32(1)
            wts = ((x1f - ix) * (y1f - iy), (ix - x0f) * (y1f - iy),
                   (x1f - ix) * (iy - y0f), (ix - x0f) * (iy - y0f))
            # clip before the int cast so far-away points cannot overflow
            x0 = np.clip(x0f, -2, w + 1).astype(np.int64)
            y0 = np.clip(y0f, -2, h + 1).astype(np.int64)
            taps = ((x0, y0), (x0 + 1, y0), (x0, y0 + 1), (x0 + 1, y0 + 1))
            img = planes[i, p].transpose(1, 2, 0)                              # [H,W,C]
            acc = np.zeros((xyz.shape[1], c), F32)
            for (tx, ty), wt in zip(taps, wts):
                ok = (tx >= 0) & (tx < w) & (ty >= 0) & (ty < h)
                v = img[np.clip(ty, 0, h - 1), np.clip(tx, 0, w - 1)]
                acc = acc + np.where(ok[:, None], v * wt.astype(F32)[:, None], F32(0))
            out[i, p] = acc
    return out


# --------------------------------------------------------------------------
# a5/a6. OSGDecoder + FullyConnectedLayer
#        training/triplane.py:113-136, training/networks_stylegan2.py:103-134
# --------------------------------------------------------------------------
@dataclass
class DecoderParams:
    """Raw OSGDecoder parameters (net.0 / net.2) and the lr multiplier."""
    w1: np.ndarray           # [64, 32]  net.0.weight
    b1: np.ndarray           # [64]      net.0.bias
    w2: np.ndarray           # [33, 64]  net.2.weight
    b2: np.ndarray           # [33]      net.2.bias
    lr_mul: float = 1.0

    def effective(self):
        """Weights/biases with the runtime gains applied
        (networks_stylegan2.py:118-119,122-127): w*lr_mul/sqrt(fan_in), b*lr_mul."""
        g1 = F32(self.lr_mul / np.sqrt(self.w1.shape[1]))
        g2 = F32(self.lr_mul / np.sqrt(self.w2.shape[1]))
        bg = F32(self.lr_mul)
        return (self.w1.astype(F32) * g1, self.b1.astype(F32) * bg,
                self.w2.astype(F32) * g2, self.b2.astype(F32) * bg)


def make_decoder_params(rng: np.random.RandomState, lr_mul: float = 1.0,
                        bias_scale: float = 0.0) -> DecoderParams:
    """Default-init decoder (randn/lr_mul weights, zero bias,
    networks_stylegan2.py:116-117); ``bias_scale`` > 0 adds non-trivial biases
    so tests also exercise the bias path."""
    w1 = (rng.standard_normal((64, 32)) / lr_mul).astype(F32)
    w2 = (rng.standard_normal((33, 64)) / lr_mul).astype(F32)
    b1 = (rng.standard_normal(64) * bias_scale).astype(F32)
    b2 = (rng.standard_normal(33) * bias_scale).astype(F32)
    return DecoderParams(w1, b1, w2, b2, lr_mul)


def softplus(x: np.ndarray) -> np.ndarray:
    """torch Softplus(beta=1, threshold=20): log1p(exp(x)), identity above 20."""
    x = np.asarray(x, F32)
    with np.errstate(over='ignore'):
        y = np.log1p(np.exp(np.minimum(x, F32(20)))).astype(F32)
    return np.where(x > F32(20), x, y).astype(F32)


def decode(features: np.ndarray, dec: DecoderParams):
    """features [N,3,P,32] -> rgb [N,P,32], sigma [N,P,1].

    mean over the three planes (triplane.py:126), FC 32->64, softplus, FC 64->33
    (:118-122), rgb = sigmoid(x[1:])*1.002-0.001 (:134), sigma = x[0] (:135).
    """
    w1, b1, w2, b2 = dec.effective()
    f = np.asarray(features, F32)
    x = ((f[:, 0] + f[:, 1] + f[:, 2]) / F32(3)).astype(F32)
    n, p, c = x.shape
    x = x.reshape(n * p, c)
    hid = softplus((x @ w1.T).astype(F32) + b1)
    out = ((hid @ w2.T).astype(F32) + b2).reshape(n, p, -1)
    with np.errstate(over='ignore'):
        sig = (F32(1) / (F32(1) + np.exp(-out[..., 1:]))).astype(F32)
    rgb = sig * F32(1 + 2 * 0.001) - F32(0.001)
    return rgb.astype(F32), out[..., 0:1].astype(F32)


def run_model(planes, dec: DecoderParams, xyz, box_warp: float, density_noise: float = 0.0, sigma_noise=None):
    """a8: ImportanceRenderer.run_model (VR/renderer.py:142-148).  ``sigma_noise`` [N,P,1] stands for the
    torch.randn_like draw of :146 (only read when density_noise > 0)."""
    rgb, sigma = decode(gather_planes(planes, xyz, box_warp), dec)
    if density_noise > 0:
        sigma = (sigma + np.asarray(sigma_noise, F32).reshape(sigma.shape) * F32(density_noise)).astype(F32)
    return rgb, sigma


# --------------------------------------------------------------------------
# a14. get_ray_limits_box                                 VR/math_utils.py:46-98
# --------------------------------------------------------------------------
def ray_limits_box(origins: np.ndarray, dirs: np.ndarray, box_side_length: float):
    """origins / dirs [...,3] -> (t_min [...,1], t_max [...,1]): slab test against the cube of side
    ``box_side_length`` centred on the origin, x then y then z; rays that miss get (-1, -2)."""
    o = np.asarray(origins, F32).reshape(-1, 3)
    d = np.asarray(dirs, F32).reshape(-1, 3)
    half = F32(box_side_length / 2)
    with np.errstate(divide='ignore', invalid='ignore'):
        inv = (F32(1) / d).astype(F32)
        neg = inv < 0
        near = np.where(neg, half, -half).astype(F32)          # the bound hit first along each axis
        far = np.where(neg, -half, half).astype(F32)
        t0 = ((near - o) * inv).astype(F32)
        t1 = ((far - o) * inv).astype(F32)
        valid = np.ones(o.shape[0], bool)
        tmin, tmax = t0[:, 0], t1[:, 0]
        for ax in (1, 2):
            valid &= ~((tmin > t1[:, ax]) | (t0[:, ax] > tmax))
            tmin, tmax = np.maximum(tmin, t0[:, ax]), np.minimum(tmax, t1[:, ax])
    tmin = np.where(valid, tmin, F32(-1)).astype(F32)
    tmax = np.where(valid, tmax, F32(-2)).astype(F32)
    shape = tuple(np.asarray(origins).shape[:-1]) + (1,)
    return tmin.reshape(shape), tmax.reshape(shape)


def auto_ray_limits(origins: np.ndarray, dirs: np.ndarray, box_warp: float):
    """The 'auto' branch of forward (VR/renderer.py:91-96): box limits per ray; rays that miss the box get
    ray_start = min and ray_end = MAX OF THE VALID ray_starts (sic, :95-96).  -> ([N,M,1], [N,M,1])."""
    rs, re = ray_limits_box(origins, dirs, box_warp)
    valid = re > rs
    if valid.any():
        lo, hi = rs[valid].min(), rs[valid].max()
        rs = np.where(valid, rs, lo).astype(F32)
        re = np.where(valid, re, hi).astype(F32)
    return rs, re


# --------------------------------------------------------------------------
# a15. sample_from_3dgrid (dead code in the reference)    VR/renderer.py:67-80
# --------------------------------------------------------------------------
def sample_from_3dgrid(grid: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """grid [1 or N,C,D,H,W], coords [N,P,3] (x -> W, y -> H, z -> D) -> [N,P,C]: trilinear, zeros padding,
    align_corners=False (5-D grid_sample)."""
    grid, coords = np.asarray(grid, F32), np.asarray(coords, F32)
    n, p, _ = coords.shape
    g, c, dd, hh, ww = grid.shape
    out = np.zeros((n, p, c), np.float64)
    pos = [((coords[..., k] + 1) * size - 1) / 2 for k, size in enumerate((ww, hh, dd))]
    base = [np.floor(q) for q in pos]
    for corner in range(8):
        off = [(corner >> k) & 1 for k in range(3)]
        idx = [b.astype(np.int64) + o for b, o in zip(base, off)]
        wt = np.ones((n, p), np.float64)
        for q, b, o in zip(pos, base, off):
            wt *= (q - b) if o else (1 - (q - b))
        ok = np.ones((n, p), bool)
        for i, size in zip(idx, (ww, hh, dd)):
            ok &= (i >= 0) & (i < size)
        xi, yi, zi = (np.clip(i, 0, size - 1) for i, size in zip(idx, (ww, hh, dd)))
        for b in range(n):
            v = grid[0 if g == 1 else b][:, zi[b], yi[b], xi[b]].T          # [P,C]
            out[b] += np.where(ok[b][:, None], v * wt[b][:, None], 0.0)
    return out.astype(F32)


# --------------------------------------------------------------------------
# a7. sample_stratified                                   VR/renderer.py:169-192
# --------------------------------------------------------------------------
def torch_linspace(start: float, end: float, steps: int) -> np.ndarray:
    """float32 torch.linspace: start+i*step for the lower half, end-(steps-1-i)*step
    for the upper half (ATen RangeFactories), step computed in float32.  This is the per-element
    formula of the CUDA kernel (and of the CPU kernel's scalar tail); the CPU kernel's vectorised
    body computes base+lane*step per SIMD vector instead, which differs from it by <= 2 ulp
    (tests/golden/stratified.npz, generated on CPU, pins that bound)."""
    start, end = F32(start), F32(end)
    step = (end - start) / F32(steps - 1)
    i = np.arange(steps)
    lo = start + step * i.astype(F32)
    hi = end - step * (steps - 1 - i).astype(F32)
    return np.where(i < steps // 2, lo, hi).astype(F32)


def stratified_depths(jitter: np.ndarray, ray_start: float, ray_end: float,
                      disparity: bool = False) -> np.ndarray:
    """jitter [N,M,D,1] in [0,1) -> coarse depths [N,M,D,1].

    Scalar-limits branch (VR/renderer.py:188-190) and the disparity branch
    (:174-181).  ``jitter`` stands for the reference's torch.rand_like draw.
    """
    jitter = np.asarray(jitter, F32)
    d = jitter.shape[2]
    if disparity:
        t = torch_linspace(0.0, 1.0, d).reshape(1, 1, d, 1) + jitter * F32(1.0 / (d - 1))
        return (F32(1) / (F32(1.0 / ray_start) * (F32(1) - t) + F32(1.0 / ray_end) * t)).astype(F32)
    base = torch_linspace(ray_start, ray_end, d).reshape(1, 1, d, 1)
    return (base + jitter * F32((ray_end - ray_start) / (d - 1))).astype(F32)


def stratified_depths_per_ray(jitter: np.ndarray, ray_start: np.ndarray, ray_end: np.ndarray) -> np.ndarray:
    """Tensor-limits branch (VR/renderer.py:183-186 with math_utils.linspace, VR/math_utils.py:101-118):
    jitter [N,M,D,1], ray_start / ray_end [N,M,1] -> [N,M,D,1]."""
    jitter = np.asarray(jitter, F32)
    rs, re = np.asarray(ray_start, F32), np.asarray(ray_end, F32)
    d = jitter.shape[2]
    steps = (np.arange(d, dtype=F32) / F32(d - 1)).reshape(1, 1, d, 1)
    base = rs[:, :, None] + steps * (re - rs)[:, :, None]
    delta = (re - rs) / F32(d - 1)
    return (base + jitter * delta[:, :, None]).astype(F32)


# --------------------------------------------------------------------------
# a9. MipRayMarcher2.run_forward                          VR/ray_marcher.py:25-57
# --------------------------------------------------------------------------
def march(colors, densities, depths, white_back: bool = False, depth_range=None):
    """colors [N,M,S,C], densities [N,M,S,1], depths [N,M,S,1] ->
    (rgb [N,M,C], depth [N,M,1], weights [N,M,S-1,1]).

    Midpoint rule (:26-29), softplus(sigma-1) (:33), alpha = 1-exp(-sigma*delta)
    (:39), transmittance = exclusive cumprod of (1-alpha+1e-10) (:41-42),
    depth = sum(w*d)/sum(w) with NaN -> inf and a clamp to the min/max over the
    WHOLE depths tensor (:46-50), white_back (:52-53), rgb*2-1 (:55).
    ``depth_range`` overrides the global (min,max) when a caller shards rays.
    """
    colors, densities, depths = (np.asarray(a, F32) for a in (colors, densities, depths))
    deltas = depths[:, :, 1:] - depths[:, :, :-1]
    c_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / F32(2)
    s_mid = softplus((densities[:, :, :-1] + densities[:, :, 1:]) / F32(2) - F32(1))
    d_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / F32(2)
    alpha = F32(1) - np.exp(-(s_mid * deltas)).astype(F32)
    keep = (F32(1) - alpha + F32(1e-10)).astype(F32)
    trans = np.ones_like(alpha)
    # sequential float32 product, like torch's CPU cumprod
    for i in range(1, alpha.shape[2]):
        trans[:, :, i] = trans[:, :, i - 1] * keep[:, :, i - 1]
    weights = (alpha * trans).astype(F32)
    rgb = (weights * c_mid).sum(2, dtype=F32)
    w_tot = weights.sum(2, dtype=F32)
    with np.errstate(invalid='ignore', divide='ignore'):
        depth = (weights * d_mid).sum(2, dtype=F32) / w_tot
    depth = np.where(np.isnan(depth), F32(np.inf), depth)
    lo, hi = depth_range if depth_range is not None else (depths.min(), depths.max())
    depth = np.clip(depth, F32(lo), F32(hi)).astype(F32)
    if white_back:
        rgb = rgb + F32(1) - w_tot
    rgb = rgb * F32(2) - F32(1)
    return rgb.astype(F32), depth, weights


# --------------------------------------------------------------------------
# a10/a11. sample_importance + sample_pdf                 VR/renderer.py:194-253
# --------------------------------------------------------------------------
def smooth_weights(weights: np.ndarray) -> np.ndarray:
    """[R,S-1] -> [R,S-1]: max_pool1d(2,1,pad=1) then avg_pool1d(2,1) then +0.01
    (VR/renderer.py:205-207)."""
    w = np.asarray(weights, F32)
    ninf = np.full((w.shape[0], 1), -np.inf, F32)
    padded = np.concatenate([ninf, w, ninf], 1)
    mp = np.maximum(padded[:, :-1], padded[:, 1:])              # S entries
    return ((mp[:, :-1] + mp[:, 1:]) / F32(2) + F32(0.01)).astype(F32)


def pdf_to_cdf(weights: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """[R,B] -> cdf [R,B+1] with a leading zero (VR/renderer.py:227-230);
    see the arithmetic contract in the module docstring."""
    w = (np.asarray(weights, F32) + F32(eps)).astype(F32)
    tot = w.astype(np.float64).sum(-1, keepdims=True).astype(F32)
    pdf = (w / tot).astype(F32)
    cdf = np.cumsum(pdf.astype(np.float64), -1).astype(F32)
    return np.concatenate([np.zeros_like(cdf[:, :1]), cdf], -1)


def sample_pdf(bins: np.ndarray, weights: np.ndarray, u: np.ndarray, eps: float = 1e-5):
    """bins [R,B+1], weights [R,B], u [R,K] -> (samples [R,K], inds [R,K] int64).

    Inverse-CDF sampling (VR/renderer.py:240-252): inds = searchsorted(cdf, u,
    right=True); below = max(inds-1,0); above = min(inds,B); a degenerate
    interval (cdf difference < eps) gets denominator 1.
    """
    bins = np.asarray(bins, F32)
    u = np.asarray(u, F32)
    cdf = pdf_to_cdf(weights, eps)
    nb = weights.shape[1]
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, nb)
    cdf_b, cdf_a = np.take_along_axis(cdf, below, 1), np.take_along_axis(cdf, above, 1)
    bin_b, bin_a = np.take_along_axis(bins, below, 1), np.take_along_axis(bins, above, 1)
    denom = cdf_a - cdf_b
    denom = np.where(denom < F32(eps), F32(1), denom).astype(F32)
    samples = bin_b + (u - cdf_b) / denom * (bin_a - bin_b)
    return samples.astype(F32), inds


def sample_importance(depths: np.ndarray, weights: np.ndarray, u: np.ndarray):
    """depths [N,M,S,1], coarse weights [N,M,S-1,1], u [N*M,K] ->
    (fine depths [N,M,K,1], inds [N*M,K]) (VR/renderer.py:194-212): bins are the
    depth midpoints (:209) and the pdf is the smoothed weights minus both ends (:210)."""
    n, m, s, _ = depths.shape
    z = np.asarray(depths, F32).reshape(n * m, s)
    w = smooth_weights(np.asarray(weights, F32).reshape(n * m, s - 1))
    z_mid = F32(0.5) * (z[:, :-1] + z[:, 1:])
    samples, inds = sample_pdf(z_mid, w[:, 1:-1], u)
    return samples.reshape(n, m, -1, 1), inds


# --------------------------------------------------------------------------
# a12. unify_samples                                      VR/renderer.py:157-167
# --------------------------------------------------------------------------
def unify_samples(d1, c1, s1, d2, c2, s2):
    """Concatenate coarse and fine samples and sort every ray by depth."""
    d = np.concatenate([d1, d2], 2)
    c = np.concatenate([c1, c2], 2)
    s = np.concatenate([s1, s2], 2)
    order = np.argsort(d, axis=2, kind='stable')
    return (np.take_along_axis(d, order, 2), np.take_along_axis(c, order, 2),
            np.take_along_axis(s, order, 2))


# --------------------------------------------------------------------------
# a13. ImportanceRenderer.forward                          VR/renderer.py:88-140
# --------------------------------------------------------------------------
def render(planes, dec: DecoderParams, origins, dirs, options: dict,
           jitter: np.ndarray, u: np.ndarray, return_stages: bool = False, density_noise_draws=None):
    """Full forward with the two random draws supplied by the caller:
    ``jitter`` [N,M,Dc,1] stands for torch.rand_like at VR/renderer.py:190 and
    ``u`` [N*M,Df] for torch.rand at :237.  ray_start == ray_end == 'auto' takes the
    per-ray box limits (:91-97).  options['density_noise'] > 0 (:146) reads
    ``density_noise_draws`` = (coarse [N,M*Dc,1], fine [N,M*Df,1]), the torch.randn_like draws.

    Returns (rgb [N,M,32], depth [N,M,1], weight_sum [N,M,1]) like :140.
    """
    origins, dirs = np.asarray(origins, F32), np.asarray(dirs, F32)
    n, m, _ = origins.shape
    bw = options['box_warp']
    white = bool(options.get('white_back', False))
    dn = float(options.get('density_noise', 0) or 0)
    nz_c, nz_f = density_noise_draws if dn > 0 else (None, None)
    assert options.get('clamp_mode', 'softplus') == 'softplus'       # ray_marcher.py:32-35
    if isinstance(options['ray_start'], str):
        assert options['ray_start'] == options['ray_end'] == 'auto'
        rs, re = auto_ray_limits(origins, dirs, bw)
        d_c = stratified_depths_per_ray(jitter, rs, re)
    else:
        d_c = stratified_depths(jitter, options['ray_start'], options['ray_end'],
                                options.get('disparity_space_sampling', False))
    dc = d_c.shape[2]
    xyz = (origins[:, :, None, :] + d_c * dirs[:, :, None, :]).reshape(n, -1, 3)
    rgb_c, sig_c = run_model(planes, dec, xyz, bw, dn, nz_c)
    rgb_c, sig_c = rgb_c.reshape(n, m, dc, -1), sig_c.reshape(n, m, dc, 1)
    stages = {'depths_coarse': d_c, 'rgb_coarse': rgb_c, 'sigma_coarse': sig_c}
    df = int(options.get('depth_resolution_importance', 0))
    if df > 0:
        _, _, w_c = march(rgb_c, sig_c, d_c, white)
        d_f, inds = sample_importance(d_c, w_c, u)
        xyz = (origins[:, :, None, :] + d_f * dirs[:, :, None, :]).reshape(n, -1, 3)
        rgb_f, sig_f = run_model(planes, dec, xyz, bw, dn, nz_f)
        rgb_f, sig_f = rgb_f.reshape(n, m, df, -1), sig_f.reshape(n, m, df, 1)
        d_all, c_all, s_all = unify_samples(d_c, rgb_c, sig_c, d_f, rgb_f, sig_f)
        rgb, depth, w = march(c_all, s_all, d_all, white)
        stages.update(weights_coarse=w_c, depths_fine=d_f, inds=inds,
                      rgb_fine=rgb_f, sigma_fine=sig_f, depths_all=d_all)
    else:
        rgb, depth, w = march(rgb_c, sig_c, d_c, white)
    out = (rgb, depth, w.sum(2, dtype=F32))
    return (out, stages) if return_stages else out


# --------------------------------------------------------------------------
# Synthetic workload shared by tests / bench / smoke
# --------------------------------------------------------------------------
FFHQ_OPTIONS = {  # train.py:312-313,328-332
    'ray_start': 2.25, 'ray_end': 3.3, 'box_warp': 1, 'depth_resolution': 48,
    'depth_resolution_importance': 48, 'disparity_space_sampling': False,
    'clamp_mode': 'softplus',
}
FFHQ_FOCAL = 4.2647       # gen_videos.py:135


def lookat_pose(h: float, v: float, radius: float) -> np.ndarray:
    """cam2world [4,4] like camera_utils.LookAtPoseSampler.sample
    (camera_utils.py:88-106: theta=h, phi=v un-remapped, camera looks at the
    origin) + create_cam2world_matrix (:155-174: y up, no roll)."""
    org = np.array([radius * math.sin(v) * math.cos(math.pi - h),
                    radius * math.cos(v),
                    radius * math.sin(v) * math.sin(math.pi - h)], np.float64)
    fwd = -org / np.linalg.norm(org)
    up = np.array([0.0, 1.0, 0.0])
    right = -np.cross(up, fwd)
    right /= np.linalg.norm(right)
    up2 = np.cross(fwd, right)
    up2 /= np.linalg.norm(up2)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up2, fwd, org
    return m.astype(F32)


def orbit_cameras(n: int, radius: float = 2.7, frames: int = 120):
    """n cameras on the gen_videos orbit (gen_videos.py:155-158; note the
    literal 3.14) at frames 0, frames/n, ...; returns (c2w [n,4,4], K [n,3,3])."""
    c2w = []
    for j in range(n):
        i = (j * frames) // max(n, 1)
        c2w.append(lookat_pose(3.14 / 2 + 0.7 * math.sin(2 * 3.14 * i / frames),
                               3.14 / 2 - 0.05 + 0.3 * math.cos(2 * 3.14 * i / frames), radius))
    K = np.array([[FFHQ_FOCAL, 0, 0.5], [0, FFHQ_FOCAL, 0.5], [0, 0, 1]], F32)
    return np.stack(c2w), np.broadcast_to(K, (n, 3, 3)).copy()


def synthetic_scene(seed: int, n_img: int, res: int, plane_res: int = 256,
                    dc: int = 48, df: int = 48, bias_scale: float = 0.0):
    """Seeded inputs of the BASELINE shape: N(0,1) planes, default-init decoder,
    orbit cameras, and the two uniform draws.  Uses numpy's frozen RandomState
    stream so fixtures regenerate bit-identically everywhere."""
    rng = np.random.RandomState(seed)
    planes = rng.standard_normal((n_img, 3, 32, plane_res, plane_res)).astype(F32)
    dec = make_decoder_params(rng, 1.0, bias_scale)
    c2w, K = orbit_cameras(n_img)
    origins, dirs = ray_sample(c2w, K, res)
    m = res * res
    below_one = np.nextafter(F32(1), F32(0))          # torch.rand never returns 1.0
    jitter = np.minimum(rng.random_sample((n_img, m, dc, 1)).astype(F32), below_one)
    u = np.minimum(rng.random_sample((n_img * m, df)).astype(F32), below_one)
    return dict(planes=planes, dec=dec, c2w=c2w, K=K, origins=origins, dirs=dirs,
                jitter=jitter, u=u)
