"""GPU setup kernels (rows f-3, f-1) against their host (numpy) specifications, which the CPU suite pins against the
reference procedures (tests/test_host_logic.py)."""
import numpy as np
import pytest
import torch

from skelsplat_b200 import configs, heatmaps, setup_gpu, synthetic, trainer, triangulation

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person-8v"])
def test_dlt_matches_numpy_svd(name):
    cfg = configs.get_config(name)
    seq = synthetic.make_sequence(cfg, 16, seed=3)
    P_list = [c.P3x4() for c in seq.cameras]
    det = np.stack([f.poses_2d for f in seq.frames])
    got = setup_gpu.triangulate_dlt(P_list, det, DEV).cpu().numpy()
    want = np.stack([triangulation.triangulate_poses(P_list, d) for d in det])
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-6            # millimetres, fp64 on both sides
    assert np.abs(want - np.stack([f.pose_3d_init for f in seq.frames])).max() < 1e-9


@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person"])
def test_heatmap_rois_match_the_host_generator(name):
    cfg = configs.get_config(name)
    seq = synthetic.make_sequence(cfg, 3, seed=7)
    seq.frames[0].poses_2d[0, 0] = [2.3, 1.1]           # corner: reflect + clamp
    seq.frames[1].poses_2d[1, 2] = [1e5, -40.0]         # outside: clamped to the border
    poses_init = np.stack([f.pose_3d_init for f in seq.frames]); poses_2d = np.stack([f.poses_2d for f in seq.frames])
    host = trainer.pack_host(cfg, seq.cameras, poses_init, poses_2d)
    ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, poses_2d, poses_init, DEV)
    rect = ps.roi_rect.cpu().numpy(); off = ps.roi_offset.cpu().numpy(); data = ps.roi_data.cpu().numpy()
    assert np.array_equal(rect, host["roi_rect"])     # integer work: every window identical (fp64 sigma, same operation order)
    assert np.array_equal(off, host["roi_offset"])
    F, V, J = rect.shape[:3]
    peak = 0.0
    for f in range(F):
        for v in range(V):
            for j in range(J):
                w, h = rect[f, v, j, 2], rect[f, v, j, 3]
                o, ho = off[f, v, j], host["roi_offset"][f, v, j]
                col, row = data[o:o + h], data[o + h:o + h + w]                       # factored patch: col[h] | row[w]
                hcol, hrow = host["roi_data"][ho:ho + h], host["roi_data"][ho + h:ho + h + w]
                assert np.abs(col - hcol).max() <= 1e-6 * hcol.max() and np.abs(row - hrow).max() <= 1e-6 * hrow.max()
                a, b = col[:, None] * row[None, :], hcol[:, None] * hrow[None, :]      # the heatmap values (one fp32 product each)
                assert np.abs(a - b).max() < 2e-6
                assert np.array_equal(a > 0, b > 0) and (a > 0).all()               # the loss mask {gt > 0} is the same set: the window
                peak = max(peak, float(a.max()))
    assert abs(peak - 1.0) < 1e-6
    assert data.min() >= 0.0


@pytest.mark.parametrize("name", ["h36m", "h36m-occ", "panoptic", "occlusion-person-8v"])
def test_heatmap_windows_identical_on_many_frames(name):
    """The GPU-generated windows (the inputs of the headline e2e path) equal the host specification's on EVERY patch of
    256 frames -- incl. anisotropic / rotated initial Gaussians, which make every term of the covariance non-trivial."""
    cfg = configs.get_config(name)
    F = 256
    seq = synthetic.make_sequence(cfg, F, seed=11)
    poses_init = np.stack([f.pose_3d_init for f in seq.frames]); poses_2d = np.stack([f.poses_2d for f in seq.frames])
    xyz, scal, rot, opa = trainer.initial_raw_state(cfg, poses_init)
    rng = np.random.default_rng(2)
    scal[F // 2:] += rng.uniform(-0.5, 0.5, scal[F // 2:].shape).astype(np.float32)
    rot[F // 2:] = rng.normal(size=rot[F // 2:].shape).astype(np.float32)
    vm, pm, dims, tanfov = trainer.camera_tensors(seq.cameras, DEV)
    Wm, Hm = max(c.image_width for c in seq.cameras), max(c.image_height for c in seq.cameras)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    rect, off, data = setup_gpu.generate_heatmap_rois_gpu(cfg, vm, pm, dims, tanfov, Wm, Hm, t(xyz), t(scal), t(rot), t(poses_2d.astype(np.float32)))
    rect = rect.cpu().numpy()
    want = np.stack([heatmaps.heatmap_roi_rects(poses_init[f], poses_2d[f].astype(np.float32), seq.cameras, scal[f], rot[f]) for f in range(F)])
    assert np.array_equal(rect, want)


def test_whole_gpu_pipeline_detections_to_poses():
    """detections -> DLT -> ROIs -> fused optimiser, all on the GPU, equals the host-prepared run (non-chaotic config)."""
    cfg = configs.OCCLUSION_PERSON
    seq = synthetic.make_sequence(cfg, 4, seed=9)
    poses_2d = np.stack([f.poses_2d for f in seq.frames])
    ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, poses_2d, None, DEV)          # initial guess from the GPU DLT
    xyz, _ = trainer.optimize_packed(ps)
    ref = trainer.optimize_sequence(seq, DEV)
    assert np.linalg.norm(xyz.cpu().numpy() - ref, axis=-1).max() < 0.05


def test_streaming_detections_in_poses_out():
    """StreamingOptimizer.submit_detections (pinned detections -> GPU DLT / initial state / heatmap ROIs -> fused optimiser ->
    poses, no host round trip) returns bit-for-bit what the exactly-sized synchronous pipeline returns, with and without
    host-provided initial poses; a streaming buffer that is too small is detected on the device and the batch is re-run
    through the synchronous path (same result)."""
    cfg = configs.H36M
    F = 6
    seqs = [synthetic.make_sequence(cfg, F, seed=50 + i) for i in range(3)]
    cams = seqs[0].cameras
    hosts, refs = [], []
    for n, sq in enumerate(seqs):
        p2 = torch.from_numpy(np.stack([f.poses_2d for f in sq.frames]).astype(np.float32)).pin_memory()
        h = {"poses_2d": p2}
        if n == 1:                                                   # this batch brings its own initial guess
            h["xyz"] = torch.from_numpy(np.stack([f.pose_3d_init for f in sq.frames]).astype(np.float32)).pin_memory()
        hosts.append(h)
        ps = setup_gpu.pack_sequence_gpu(cfg, cams, p2, h.get("xyz"), DEV)
        cap = int(ps.roi_data.numel())
        refs.append(trainer.optimize_packed(ps, iterations=40)[0].cpu().numpy())
    so = trainer.StreamingOptimizer(cfg, cams, F, 2 * cap, DEV, iterations=40)
    t = [so.submit_detections(h) for h in hosts[:2]]
    outs = [so.result(t[0])]
    t.append(so.submit_detections(hosts[2]))
    outs += [so.result(t[1]), so.result(t[2])]
    for a, b in zip(outs, refs):
        assert np.array_equal(a, b)
    assert so.launches == 3
    small = trainer.StreamingOptimizer(cfg, cams, F, cap // 2, DEV, iterations=40)      # ROI patches do not fit
    assert np.array_equal(small.result(small.submit_detections(hosts[0])), refs[0])
    assert int(small.slots[0]["det"]["setup_status_host"][0]) & 2
