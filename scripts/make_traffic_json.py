"""Developer tool: profiles/<round>_traffic.json from the ncu captures of gpu_refresh_profiles.sh -- the per-launch DRAM bytes,
executed instructions and issue-slot utilisation bench.py scales into `roofline.traffic` / `issue_roofline`.  The file records
the source hash of the library the captures were made with; bench.py ignores it when the loaded library differs.
usage: make_traffic_json.py [round] [out_dir]   (needs `ncu -i` and the .ncu-rep files under gpurun_out/)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT_DIR = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return [{h: (v, u) for h, u, v in zip(r[0], r[1], vals)} for vals in r[2:]]


UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}


def num(d, k):
    v, u = d[k]
    return float(v.replace(",", "")) * UNIT.get(u, 1.0)       # bytes, milliseconds, plain counts / percentages


from skelsplat_b200 import build
MAN = dict(e.split(":") for e in build.source_manifest().split(","))
OPT_FILES = ("optimizer.cu", "common.cuh", "adam_form.h", "api_internal.h", "skelsplat_b200.h")
DENSE_FILES = ("raster_dense.cu", "common.cuh", "api_internal.h", "skelsplat_b200.h")
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum, smsp__issue_active from `ncu --set full` captures "
                   f"(gpurun_out/{R}_*.ncu-rep, summaries in profiles/{R}_*_ncu.txt); bench.py scales them to its launch size and drops them "
                   "when the files that define the captured kernel differ in the loaded library (ssb_source_manifest)",
       "kernel_sources": {"optimize_kernel": {f: MAN[f] for f in OPT_FILES}, "dense_rasterizer": {f: MAN[f] for f in DENSE_FILES}}, "optimize_kernel": {}}
for cfg in ("h36m", "h36m-occ", "panoptic", "occlusion-person-8v"):
    rep = os.path.join(ROOT, "gpurun_out", f"{R}_opt_{cfg}.ncu-rep")
    if not os.path.exists(rep):
        continue
    d = rows(rep)[0]
    out["optimize_kernel"][cfg] = {"frames_in_capture": int(num(d, "launch__grid_size")), "dram_bytes_read": num(d, "dram__bytes_read.sum"),
                                   "dram_bytes_write": num(d, "dram__bytes_write.sum"), "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                   "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"), "inst_executed": num(d, "smsp__inst_executed.sum"),
                                   "duration_ms_under_ncu": num(d, "gpu__time_duration.sum"), "source": f"profiles/{R}_optimize_kernel_{cfg}_ncu.txt"}
rep = os.path.join(ROOT, "gpurun_out", f"{R}_dense.ncu-rep")
if os.path.exists(rep):
    rs = rows(rep)
    out["dense_rasterizer"] = {"views_in_capture": 16, "dram_bytes_read": sum(num(d, "dram__bytes_read.sum") for d in rs),
                               "dram_bytes_write": sum(num(d, "dram__bytes_write.sum") for d in rs), "source": f"profiles/{R}_dense_rasterizer_ncu.txt"}
json.dump(out, open(os.path.join(OUT_DIR, f"{R}_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
