import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gnerf_b200 as pkg
from oracle import ref_loader
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
pkg.enable_reference_plugins()
ref = ref_loader.import_reference()
dev = torch.device('cuda:0')
G = ref_loader.make_generator(seed=3, depth_resolution=96, depth_resolution_importance=96).to(dev)
intr = torch.tensor([[4.2647, 0, 0.5], [0, 4.2647, 0.5], [0, 0, 1]], device=dev)
z = torch.randn((2, 512), generator=torch.Generator().manual_seed(4)).to(dev)
pose = ref.camera_utils.LookAtPoseSampler.sample(3.14/2, 3.14/2-0.05+0.3, radius=2.7, device=dev)
c = torch.cat([pose.reshape(-1, 16), intr.reshape(-1, 9)], 1).repeat(2, 1)
with torch.no_grad():
    ws = G.mapping(z=z, c=torch.zeros_like(c))
    p1 = G.backbone.synthesis(ws, noise_mode='const')
    p2 = G.backbone.synthesis(ws, noise_mode='const')
    print('backbone run-to-run max diff', float((p1 - p2).abs().max()), 'planes std', float(p1.std()), 'absmax', float(p1.abs().max()))
    planes = p1.view(2, 3, 32, 256, 256)
    o, d = G.ray_sampler(c[:, :16].view(-1, 4, 4), c[:, 16:25].view(-1, 3, 3), 64)
    for dcf in (48, 96):
        rk = dict(G.rendering_kwargs, depth_resolution=dcf, depth_resolution_importance=dcf)
        torch.manual_seed(11); want = ref.renderer.ImportanceRenderer()(planes, G.decoder, o, d, rk)
        torch.manual_seed(11); want2 = ref.renderer.ImportanceRenderer()(planes, G.decoder, o, d, rk)
        R = pkg.ImportanceRenderer(); R.debug_outputs = True
        torch.manual_seed(11); got = R(planes, G.decoder, o, d, rk)
        for name, a, b, b2 in zip(('rgb', 'depth', 'wsum'), got, want, want2):
            e = (a - b).abs()
            print(dcf, name, 'max', float(e.max()), 'mean', float(e.mean()), 'rays>1e-4', int((e.reshape(e.shape[0], e.shape[1], -1).max(-1).values > 1e-4).sum()),
                  'ref run-to-run', float((b - b2).abs().max()), 'range', float(b.min()), float(b.max()))
        # sigma statistics through run_model
        pts = (o.unsqueeze(-2) + torch.linspace(2.25, 3.3, 48, device=dev).reshape(1, 1, 48, 1) * d.unsqueeze(-2)).reshape(2, -1, 3)
        out = R.run_model(planes, G.decoder, pts, None, rk)
        Rr = ref.renderer.ImportanceRenderer(); Rr.plane_axes = Rr.plane_axes.to(dev)
        outr = Rr.run_model(planes, G.decoder, pts, None, rk)
        print(dcf, 'run_model sigma err', float((out['sigma'] - outr['sigma']).abs().max()), 'rgb err', float((out['rgb'] - outr['rgb']).abs().max()),
              'sigma range', float(outr['sigma'].min()), float(outr['sigma'].max()))
        # bf16
        torch.manual_seed(11); g16 = R(planes, G.decoder, o, d, dict(rk, decoder_precision='bf16'))
        print(dcf, 'bf16 max err', [float((a - b).abs().max()) for a, b in zip(g16, want)])
        torch.manual_seed(11); gff = R(planes, G.decoder, o, d, dict(rk, decoder_precision='fp32_ffma'))
        print(dcf, 'ffma max err', [float((a - b).abs().max()) for a, b in zip(gff, want)])
print('decoder w stats', float(G.decoder.net[0].weight.std()), float(G.decoder.net[2].weight.std()))
