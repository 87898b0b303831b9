"""Developer tool: time launch-shape variants of the fused optimiser (threads/CTA x min CTAs/SM x r_capacity)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)

CHILD = r'''
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
import bench
from skelsplat_b200 import configs, trainer
name, F, rcap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = configs.get_config(name)
seq, host, gt = bench.make_host_batch(cfg, F, seed=100)
ps = trainer.pack_sequence(cfg, seq.cameras, host["xyz"], None, "cuda", host=host)
init = tuple(t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity))
ts = []
for rep in range(3):
    for d, s_ in zip((ps.xyz, ps.scaling, ps.rotation, ps.opacity), init): d.copy_(s_)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); trainer.optimize_packed(ps, r_capacity=rcap, check=(rep == 0)); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
x = ps.xyz.cpu().numpy()
print(json.dumps({"ms": min(ts), "fps": F / min(ts) * 1e3, "checksum": float(np.abs(x).sum()), "mpjpe": trainer.mpjpe(x, gt)}))
''' % ROOT

def main():
    from skelsplat_b200 import build
    variants = [(0, 0), (1, 0), (0, 1), (1, 1)]      # (SSB_VCULL, repeat)
    works = [("h36m", 2048, 256), ("occlusion-person-8v", 2048, 512)]
    out = {}
    for thr, cta in variants:
        lib = os.path.join(ROOT, "skelsplat_b200", "_lib", f"libvariant_{thr}_{cta}.so")
        if not os.path.exists(lib):
            build.build(force=True, defines=(f"SSB_VCULL={thr}",), out=lib)
    for name, F, rcap in works:
        for thr, cta in variants:
            lib = os.path.join(ROOT, "skelsplat_b200", "_lib", f"libvariant_{thr}_{cta}.so")
            env = dict(os.environ, SKELSPLAT_B200_LIB=lib)
            r = subprocess.run([sys.executable, "-c", CHILD, name, str(F), str(rcap)], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(name, F, "rcap", rcap, "threads", thr, "minCTAs", cta, line, flush=True)
            out[f"{name}|{rcap}|{thr}|{cta}"] = line
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_opt.json"), "w"), indent=1)

if __name__ == "__main__":
    if "--build-only" in sys.argv:
        from skelsplat_b200 import build
        for thr, cta in [(0, 0), (1, 0), (0, 1), (1, 1)]:
            build.build(force=True, defines=(f"SSB_VCULL={thr}",),
                        out=os.path.join(ROOT, "skelsplat_b200", "_lib", f"libvariant_{thr}_{cta}.so"))
    else:
        main()
