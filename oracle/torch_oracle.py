"""TEST INFRASTRUCTURE -- a torch restatement of the reference renderer's algorithm (device-agnostic).

Why a second oracle: the numpy / C oracles (triplane_oracle.py / .c) finish small cases in seconds on a CPU but cannot
check BASELINE.json's full sizes.  This one issues the same ATen operations the reference issues -- grid_sample,
addmm, softplus, cumprod, cumsum, searchsorted, sort, max_pool1d / avg_pool1d: the third-party arithmetic SURVEY.md
section 8(c) lists -- so on the GPU box it (1) checks every ray of a full config-2 image against the CUDA kernels and
(2) is the "GPU baseline" of SURVEY.md section 8(d): what the reference's eager-PyTorch renderer costs on the same B200.
/root/reference does not exist on that box, hence a restatement; it is pinned here on CPU against the reference-generated
fixtures in tests/golden/ (tests/test_torch_oracle.py).  Written from the algorithm description in SURVEY.md section 8(a)
and the numpy oracle, as one flat function per stage; nothing is copied from the reference.

Only tests/ and profiles/ scripts import this module.  Product code (g-nerf_b200/) never does.

Citations: VR/ = /root/reference/g_nerf/training/volumetric_rendering/.
"""
import torch
import torch.nn.functional as F

PLANE_UV = ((0, 1), (0, 2), (2, 0))      # plane 0 <- (x,y), 1 <- (x,z), 2 <- (z,x)   (VR/renderer.py:23-53, probed)


def gather(planes, xyz, box_warp):
    """planes [N,3,C,H,W], xyz [N,P,3] -> [N,3,P,C]: bilinear lookups on the three planes (VR/renderer.py:55-65)."""
    n, _, c, h, w = planes.shape
    g = (2.0 / box_warp) * xyz
    uv = torch.stack([g[..., list(sel)] for sel in PLANE_UV], 1).reshape(n * 3, 1, -1, 2)
    out = F.grid_sample(planes.reshape(n * 3, c, h, w), uv, mode='bilinear', padding_mode='zeros', align_corners=False)
    return out.permute(0, 3, 2, 1).reshape(n, 3, -1, c)


def decode(features, dec):
    """[N,3,P,32] -> rgb [N,P,32], sigma [N,P,1] (training/triplane.py:124-136; gains: networks_stylegan2.py:118-134).
    dec: (w1 [64,32], b1 [64], w2 [33,64], b2 [33], lr_mul)."""
    w1, b1, w2, b2, lr = dec
    x = features.mean(1)
    n, p, c = x.shape
    x = x.reshape(n * p, c)
    hid = F.softplus(torch.addmm((b1 * lr).unsqueeze(0), x, (w1 * (lr / c ** 0.5)).t()))
    y = torch.addmm((b2 * lr).unsqueeze(0), hid, (w2 * (lr / 64 ** 0.5)).t()).reshape(n, p, -1)
    return torch.sigmoid(y[..., 1:]) * (1 + 2 * 0.001) - 0.001, y[..., 0:1]


def coarse_depths(jitter, ray_start, ray_end, disparity):
    """jitter [N,M,D,1] -> depths [N,M,D,1] (VR/renderer.py:169-192, scalar limits)."""
    d = jitter.shape[2]
    if disparity:
        t = torch.linspace(0, 1, d, device=jitter.device).reshape(1, 1, d, 1) + jitter * (1 / (d - 1))
        return 1. / (1. / ray_start * (1. - t) + 1. / ray_end * t)
    base = torch.linspace(ray_start, ray_end, d, device=jitter.device).reshape(1, 1, d, 1)
    return base + jitter * ((ray_end - ray_start) / (d - 1))


def march(colors, sigma, depths, white_back=False):
    """[N,M,S,C], [N,M,S,1], [N,M,S,1] -> rgb [N,M,C], depth [N,M,1], weights [N,M,S-1,1] (VR/ray_marcher.py:25-57)."""
    delta = depths[:, :, 1:] - depths[:, :, :-1]
    c_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / 2
    s_mid = F.softplus((sigma[:, :, :-1] + sigma[:, :, 1:]) / 2 - 1)
    d_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / 2
    alpha = 1 - torch.exp(-(s_mid * delta))
    shifted = torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2)
    weights = alpha * torch.cumprod(shifted, -2)[:, :, :-1]
    rgb = torch.sum(weights * c_mid, -2)
    total = weights.sum(2)
    depth = torch.sum(weights * d_mid, -2) / total
    depth = torch.nan_to_num(depth, float('inf'))
    depth = torch.clamp(depth, torch.min(depths), torch.max(depths))
    if white_back:
        rgb = rgb + 1 - total
    return rgb * 2 - 1, depth, weights


def importance_depths(depths, weights, u):
    """coarse depths [N,M,S,1], coarse weights [N,M,S-1,1], u [N*M,K] -> (fine depths [N,M,K,1], inds [N*M,K])
    (VR/renderer.py:194-253)."""
    n, m, s, _ = depths.shape
    z = depths.reshape(n * m, s)
    w = weights.reshape(n * m, 1, s - 1)
    w = F.avg_pool1d(F.max_pool1d(w, 2, 1, padding=1), 2, 1).reshape(n * m, s - 1) + 0.01
    bins = 0.5 * (z[:, :-1] + z[:, 1:])
    w = w[:, 1:-1] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below, above = torch.clamp_min(inds - 1, 0), torch.clamp_max(inds, w.shape[1])
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = ca - cb
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return (bb + (u - cb) / denom * (ba - bb)).reshape(n, m, -1, 1), inds


def render(planes, dec, origins, dirs, options, jitter, u, return_stages=False):
    """ImportanceRenderer.forward (VR/renderer.py:88-140) for scalar ray limits.
    planes [N,3,32,H,W], origins / dirs [N,M,3], jitter [N,M,Dc,1], u [N*M,Df] -> rgb [N,M,32], depth, weight sum."""
    n, m, _ = origins.shape
    dc, df = options['depth_resolution'], options['depth_resolution_importance']
    box, white = options['box_warp'], bool(options.get('white_back', False))
    d_c = coarse_depths(jitter.reshape(n, m, dc, 1), options['ray_start'], options['ray_end'],
                        bool(options.get('disparity_space_sampling', False)))

    def shade(depths):
        k = depths.shape[2]
        pts = (origins.unsqueeze(-2) + depths * dirs.unsqueeze(-2)).reshape(n, -1, 3)
        rgb, sigma = decode(gather(planes, pts, box), dec)
        return rgb.reshape(n, m, k, -1), sigma.reshape(n, m, k, 1)

    c_c, s_c = shade(d_c)
    stages = {}
    if df > 0:
        _, _, w_c = march(c_c, s_c, d_c, white)
        d_f, inds = importance_depths(d_c, w_c, u)
        c_f, s_f = shade(d_f)
        d_all, order = torch.sort(torch.cat([d_c, d_f], -2), dim=-2)
        c_all = torch.gather(torch.cat([c_c, c_f], -2), -2, order.expand(-1, -1, -1, c_c.shape[-1]))
        s_all = torch.gather(torch.cat([s_c, s_f], -2), -2, order)
        rgb, depth, w = march(c_all, s_all, d_all, white)
        stages = {'weights_coarse': w_c, 'depths_fine': d_f, 'inds': inds}
    else:
        rgb, depth, w = march(c_c, s_c, d_c, white)
    out = (rgb, depth, w.sum(2))
    return (out, stages) if return_stages else out


def decoder_tuple(dec, device):
    """oracle.triplane_oracle.DecoderParams -> the tuple decode() takes."""
    t = lambda a: torch.from_numpy(a).to(device)
    return (t(dec.w1), t(dec.b1), t(dec.w2), t(dec.b2), float(dec.lr_mul))


def render_grads(planes, dec, origins, dirs, options, jitter, u, g_rgb, g_depth, g_wsum):
    """Gradients of L = sum(rgb*g_rgb) + sum(depth*g_depth) + sum(wsum*g_wsum) w.r.t. planes and the four decoder
    tensors, by autograd through render().  The importance resampling is cut out of the graph exactly where the
    reference cuts it (torch.no_grad + detach, VR/renderer.py:198,210): importance_depths' output is detached.
    Returns ((rgb, depth, wsum), (g_planes, g_w1, g_b1, g_w2, g_b2))."""
    planes = planes.detach().clone().requires_grad_(True)
    w1, b1, w2, b2 = (p.detach().clone().requires_grad_(True) for p in dec[:4])
    global importance_depths
    real = importance_depths

    def detached(depths, weights, uu):
        with torch.no_grad():
            return real(depths, weights, uu)
    importance_depths = detached
    try:
        rgb, depth, wsum = render(planes, (w1, b1, w2, b2, dec[4]), origins, dirs, options, jitter, u)
    finally:
        importance_depths = real
    loss = (rgb * g_rgb).sum() + (depth * g_depth).sum() + (wsum * g_wsum).sum()
    grads = torch.autograd.grad(loss, (planes, w1, b1, w2, b2))
    return (rgb.detach(), depth.detach(), wsum.detach()), grads
