"""Developer tool: the worst joint of the fused optimiser vs a golden (e.g. occlusion-person-8v) -- is the reference itself
reproducible on that frame?  Runs the restated loop on the reference kernels 3x on the worst frame."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import pipeline as opipe
from skelsplat_b200 import configs, synthetic, trainer, heatmaps
from skelsplat_b200.cameras import cameras_extent
name = sys.argv[1] if len(sys.argv) > 1 else "occlusion-person-8v"
G = np.load(os.path.join(ROOT, "tests", "golden", f"opt_{name}.npz"))
cfg = configs.get_config(name)
seq = synthetic.make_sequence(cfg, int(G["n_frames"]), seed=int(G["seed"]))
mine = trainer.optimize_sequence(seq, "cuda")
dev = np.linalg.norm(mine - G["ref_xyz"], axis=-1)
f, j = np.unravel_index(dev.argmax(), dev.shape)
print("worst", name, "frame", f, "joint", j, "dev", dev[f, j], "frame devs", np.round(dev[f], 4))
fr = seq.frames[f]
_, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0])
dense = [torch.from_numpy(heatmaps.rois_to_dense(rois, v)).to("cuda") for v in range(cfg.nviews)]
runs = [opipe.optimise_frame(fr, seq.cameras, cfg, cameras_extent(seq.cameras), dense, backend="ref", device="cuda") for _ in range(3)]
for r in runs:
    print("ref rerun vs golden at worst joint", np.linalg.norm(r[j] - G["ref_xyz"][f, j]), "| vs fused", np.linalg.norm(r[j] - mine[f, j]),
          "| max over joints vs golden", np.linalg.norm(r - G["ref_xyz"][f], axis=-1).max())
