"""Developer tool: which per-Gaussian state fields differ bitwise from the reference, with examples."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from skelsplat_b200 import configs, synthetic
from skelsplat_b200 import rasterizer as R
from oracle import ref_rasterizer as refr, rast as crast
DEV = "cuda"
cfg = configs.H36M
rng = np.random.default_rng(123)
seq = synthetic.make_sequence(cfg, 1, seed=5)
cam = seq.cameras[0]; W, H = cam.image_width, cam.image_height
P = 256
means = torch.from_numpy((rng.uniform(-1500, 1500, (P, 3)) + np.array([0, 0, 900])).astype(np.float32)).to(DEV)
scales = torch.from_numpy(np.exp(rng.uniform(0.5, 4.5, (P, 3))).astype(np.float32)).to(DEV)
rots = torch.nn.functional.normalize(torch.from_numpy(rng.normal(size=(P, 4)).astype(np.float32)).to(DEV))
opac = torch.from_numpy(rng.uniform(0.05, 1.0, (P, 1)).astype(np.float32)).to(DEV)
feats = torch.from_numpy(rng.uniform(0, 1, (P, 1, 17)).astype(np.float32)).to(DEV)
vm = torch.from_numpy(cam.world_view_transform).to(DEV); pm = torch.from_numpy(cam.full_proj_transform).to(DEV)
cp = torch.from_numpy(cam.camera_center).to(DEV); e = torch.Tensor([]); bg = torch.zeros(32, device=DEV)
Rn, rcolor, rradii, geom, binning, img, rinvd = refr.rasterize_forward("h36m", bg, means, e, opac, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, feats, 0, cp, r_capacity=1 << 16)
rs = refr.RefState(geom, binning, img, Rn, P, W, H, "h36m").parse()
color, radii, invd, st = R.rasterize_batched(means[None], scales[None], rots[None], opac.reshape(1, -1), feats.reshape(P, 17), vm.reshape(1, 4, 4), pm.reshape(1, 4, 4), W, H, cam.tanfovx, cam.tanfovy, r_capacity=8192)
torch.cuda.synchronize()
ms = st.parse(0)
of = crast.forward(means.cpu().numpy(), scales.cpu().numpy(), rots.cpu().numpy(), opac.cpu().numpy(), feats.reshape(P, 17).cpu().numpy(), cam.world_view_transform, cam.full_proj_transform, W, H, cam.tanfovx, cam.tanfovy)
vis = rradii.cpu().numpy() > 0
print("visible", vis.sum(), "of", P)
for k in ("depths", "means2D", "conic_opacity", "cov3D"):
    a = ms[k][vis].reshape(vis.sum(), -1); b = rs[k][vis].reshape(vis.sum(), -1); c = of[k][vis].reshape(vis.sum(), -1)
    neq = (a.view(np.uint32) != b.view(np.uint32)); neq_o = (c.view(np.uint32) != b.view(np.uint32))
    print(k, "mine!=ref per column:", neq.sum(0), " oracle!=ref per column:", neq_o.sum(0), "mine!=oracle", (a.view(np.uint32) != c.view(np.uint32)).sum())
    idx = np.argwhere(neq)
    for (i, j) in idx[:4]:
        print("   ex", i, j, "mine", a[i, j], hex(a.view(np.uint32)[i, j]), "ref", b[i, j], hex(b.view(np.uint32)[i, j]), "oracle", c[i, j])
for k in ("tiles_touched", "point_offsets"):
    print(k, (ms[k] != rs[k]).sum())
