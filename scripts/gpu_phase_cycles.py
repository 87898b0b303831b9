"""Developer tool: per-phase SM-cycle breakdown of the fused optimiser (build with -DSSB_PHASE_TIMING=1).
   here:   python scripts/gpu_phase_cycles.py --build-only
   GPU:    python scripts/gpu_phase_cycles.py [config] [frames]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "skelsplat_b200", "_lib", "libvariant_phase.so")
NAMES = ["A activations+projection", "B scan/rank/placement", "(unused)", "B tile-run compaction", "C tiles", "D chain",
         "tail", "E Adam"]

if "--build-only" in sys.argv:
    from skelsplat_b200 import build
    build.build(force=True, defines=("SSB_PHASE_TIMING=1",), out=LIB)
    sys.exit(0)

os.environ["SKELSPLAT_B200_LIB"] = LIB
import numpy as np, torch
import bench
from skelsplat_b200 import configs, trainer, lib as L_

args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args else "h36m"
F = int(args[1]) if len(args) > 1 else 2048
cfg = configs.get_config(name)
from skelsplat_b200 import setup_gpu
seq, p2d, init0, gt = bench.make_detection_batch(cfg, F, 0)
ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, torch.from_numpy(p2d), torch.from_numpy(init0), "cuda")
lib = L_.lib()
out = (C.c_ulonglong * 8)()
trainer.optimize_packed(ps)
torch.cuda.synchronize()
lib.ssb_debug_phase_cycles(out, 0)
tot = float(sum(out))
hist = (C.c_ulonglong * 24)()
lib.ssb_debug_list_hist(hist)
ht = float(sum(hist)); he = float(sum(i * hist[i] for i in range(24)))
print(json.dumps({"tiles_by_list_length": {i: round(hist[i] / ht, 4) for i in range(24) if hist[i]},
                  "pairs_by_list_length": {i: round(i * hist[i] / he, 4) for i in range(24) if hist[i]}}))
print(json.dumps({"config": name, "frames": F, "share": {n: round(out[i] / tot, 4) for i, n in enumerate(NAMES)},
                  "cycles_per_frame_step": {n: round(out[i] / F / (cfg.iterations // cfg.accumulation_steps)) for i, n in enumerate(NAMES)}}))
