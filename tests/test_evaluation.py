"""Result sink (row f-4): MPJPE definitions of eval.py and the PLY / npz writers (host logic, CPU)."""
import numpy as np

from skelsplat_b200 import evaluation as ev


def test_mpjpe_definitions():
    rng = np.random.default_rng(0)
    gt = rng.normal(size=(5, 17, 3)) * 300
    pred = gt + rng.normal(size=gt.shape) * 10
    a = ev.mpjpe_absolute(pred, gt)
    assert a.shape == (5,)
    assert np.allclose(a[2], np.mean([np.linalg.norm(pred[2, j] - gt[2, j]) for j in range(17)]))     # eval.py:122-123
    shifted = pred + np.array([100.0, -50.0, 20.0])
    assert np.allclose(ev.mpjpe_root_relative(shifted, gt), ev.mpjpe_root_relative(pred, gt))         # translation invariant
    assert ev.mpjpe_absolute(shifted, gt).mean() > a.mean()
    names = ["S9_Walking_0", "S9_Walking_64", "S9_Eating_0", "S11_Eating_64", "S11_Eating_128"]
    rep = ev.evaluate(pred, gt, names)
    assert rep["n_frames"] == 5 and set(rep["per_action"]) == {"Walking", "Eating"} and rep["per_action"]["Eating"]["n"] == 3
    assert np.isclose(rep["absolute_mpjpe_mm"], a.mean())


def test_ply_and_npz_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    xyz = rng.normal(size=(17, 3)).astype(np.float32) * 500
    ev.write_ply(str(tmp_path / "a" / "cloud.ply"), xyz)                                              # triangulation.py style
    assert np.array_equal(ev.read_ply_xyz(str(tmp_path / "a" / "cloud.ply")), xyz)
    ev.write_ply(str(tmp_path / "full.ply"), xyz, features_dc=np.eye(17), opacity=np.full(17, np.inf), scaling=np.full((17, 3), 3.0),
                 rotation=np.tile([1, 0, 0, 0], (17, 1)))
    raw = open(tmp_path / "full.ply", "rb").read()
    header = raw[:raw.index(b"end_header")].decode()
    assert "property float f_dc_16" in header and "property float rot_3" in header and "element vertex 17" in header
    assert np.array_equal(ev.read_ply_xyz(str(tmp_path / "full.ply")), xyz)
    ev.save_poses_npz(str(tmp_path / "out" / "poses.npz"), xyz[None], ["S1_A_0"], xyz[None])
    z = np.load(tmp_path / "out" / "poses.npz")
    assert np.array_equal(z["xyz"][0], xyz) and z["scene_names"][0] == "S1_A_0"
