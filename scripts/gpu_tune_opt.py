"""Developer tool: A/B-time compile-time variants (-D defines) of the fused optimiser; prints fps, checksum, MPJPE."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)

CHILD = r'''
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
import bench
from skelsplat_b200 import configs, trainer
name, F, rcap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = configs.get_config(name)
from skelsplat_b200 import setup_gpu
seq, p2d, init0, gt = bench.make_detection_batch(cfg, F, 0)
ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, torch.from_numpy(p2d), torch.from_numpy(init0), "cuda")
init = tuple(t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity))
ts = []
for rep in range(3):
    for d, s_ in zip((ps.xyz, ps.scaling, ps.rotation, ps.opacity), init): d.copy_(s_)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        trainer.optimize_packed(ps, r_capacity=rcap, check=(rep == 0))
    except Exception as exc:      # capacity overflow with an explicit r_capacity: report how many frames
        print(json.dumps({"error": str(exc)[:160]})); sys.exit(0)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
x = ps.xyz.cpu().numpy()
print(json.dumps({"ms": min(ts), "fps": F / min(ts) * 1e3, "checksum": float(np.abs(x).sum()), "mpjpe": trainer.mpjpe(x, gt)}))
''' % ROOT

def variants_from_argv():
    """--variants "A=1,B=2;A=0" -> [("A=1","B=2"), ("A=0",)]; default: the library's defaults, twice (run-to-run noise)."""
    for i, a in enumerate(sys.argv):
        if a == "--variants":
            return [tuple(d for d in v.split(",") if d) for v in sys.argv[i + 1].split(";")]   # "...@192": r_capacity override
    return [(), ()]


def lib_path(i):
    return os.path.join(ROOT, "skelsplat_b200", "_lib", f"libvariant_{i}.so")


def build_variants(variants):
    from skelsplat_b200 import build
    for i, defs in enumerate(variants):
        build.build(force=True, defines=tuple(d.split("@")[0] for d in defs if d.split("@")[0]), out=lib_path(i))


def main(variants):
    works = [("h36m", 2048, 320), ("occlusion-person-8v", 2048, 512), ("panoptic", 1024, 1024)]
    out = {}
    for name, F, rcap in works:
        for i, defs in enumerate(variants):
            env = dict(os.environ, SKELSPLAT_B200_LIB=lib_path(i))
            over = [d.split("@")[1] for d in defs if "@" in d]
            if over and name == "h36m":
                rcap = int(over[0])
            r = subprocess.run([sys.executable, "-c", CHILD, name, str(F), str(rcap)], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(name, F, "rcap", rcap, ",".join(defs) or "default", line, flush=True)
            out[f"{name}|{rcap}|{i}|{','.join(defs)}"] = line
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_opt.json"), "w"), indent=1)


if __name__ == "__main__":
    v = variants_from_argv()
    if "--build-only" in sys.argv:      # here (no GPU): cross-compile the variants so they travel with the snapshot
        build_variants(v)
    else:
        if not all(os.path.exists(lib_path(i)) for i in range(len(v))):
            build_variants(v)
        main(v)
