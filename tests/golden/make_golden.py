"""Generate golden vectors by running the UNMODIFIED reference module.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports ``GATsSuperGlue`` from ``/root/reference`` (reference
``src/models/GATsSPG_architectures/GATs_SuperGlue.py:143``), loads the seeded
numpy state dict of ``onepose_b200.synthetic.make_state_dict`` into it, feeds
the seeded synthetic inputs of ``onepose_b200.synthetic.make_batch`` and stores
what the reference returns.  Inputs are NOT stored: they are regenerated from
the seeds recorded in each file (numpy RandomState is platform-stable).

Also stores the adjacent producers: ``mean_descriptors`` / ``mean_scores``
(reference ``src/sfm/postprocess/feature_process.py:297-317``; the module needs h5py at import
time, so the two functions are exec'd from its SOURCE TEXT, unmodified) and
``pad_features3d_random`` / ``build_features3d_leaves`` (``src/utils/data_utils.py:143-205``,
imported) under a fixed ``np.random.seed``; ``cv2.solvePnPRansac`` poses through ``ransac_PnP``
(``src/utils/eval_utils.py:18-42``); and the SuperPoint extractor (``src/models/extractors/SuperPoint/superpoint.py``,
imported) on seeded weights / images -- once as this image's torch runs it (grid_sample align_corners=False) and once with
``torch.__version__`` reading "1.8.0", the reference's pinned version, so that the module itself takes its align_corners=True
branch (superpoint.py:86).

    python tests/golden/make_golden.py [superpoint|matcher|producers|pnp ...]     (default: everything)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from onepose_b200 import synthetic  # noqa: E402
from src.models.GATsSPG_architectures.GATs_SuperGlue import GATsSuperGlue  # noqa: E402

CASES = [
    # name, weights(seed, damped), hparam overrides, object_id, frame_ids, N, M, L, conf storage
    dict(name="tiny_n64_m96", wseed=0, damped=True, hp={}, obj=1, frames=[11], N=64, M=96, L=8, full=True),
    dict(name="ragged_b2_n200_m333", wseed=0, damped=True, hp={}, obj=2, frames=[21, 22], N=200, M=333, L=8, full=True),
    dict(name="leaf3_n50_m70", wseed=1, damped=True, hp={}, obj=3, frames=[31], N=50, M=70, L=3, full=True),
    dict(name="undamped_n128_m256", wseed=2, damped=False, hp={}, obj=4, frames=[41], N=128, M=256, L=8, full=True),
    dict(name="cfg1_n512_m2048", wseed=0, damped=True, hp={}, obj=5, frames=[51], N=512, M=2048, L=8, full=False),
    dict(name="noself_n64_m96", wseed=0, damped=True, hp={"include_self": False}, obj=1, frames=[11], N=64, M=96, L=8, full=True),
    dict(name="lintrans_n64_m96", wseed=0, damped=True, hp={"with_linear_transform": True}, obj=1, frames=[11], N=64, M=96, L=8, full=True),
    dict(name="additional_n64_m96", wseed=0, damped=True, hp={"additional": True}, obj=1, frames=[11], N=64, M=96, L=8, full=True),
    dict(name="noself_lintrans_n64_m96", wseed=0, damped=True,
         hp={"include_self": False, "with_linear_transform": True}, obj=1, frames=[11], N=64, M=96, L=8, full=True),
]


def run_case(c):
    hp = dict(synthetic.DEFAULT_HPARAMS)
    hp.update(c["hp"])
    sd = synthetic.make_state_dict(c["wseed"], damped=c["damped"], hparams=hp)
    model = GATsSuperGlue(hp).eval()
    missing = model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    data = synthetic.make_batch(c["obj"], c["frames"], c["N"], c["M"], c["L"])
    inp = {k: torch.from_numpy(v) for k, v in data.items()}
    with torch.no_grad():
        pred, conf = model(inp)
        # fp64 run of the same module = accuracy referee (Tensor.float patched to identity)
        model64 = GATsSuperGlue(hp).eval().double()
        model64.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)).double() for k, v in sd.items()})
        orig_float = torch.Tensor.float
        torch.Tensor.float = lambda self, *a, **k: self
        try:
            pred64, conf64 = model64({k: v.double() for k, v in inp.items()})
        finally:
            torch.Tensor.float = orig_float
    conf = conf.numpy()
    out = {
        "meta_wseed": c["wseed"], "meta_damped": c["damped"], "meta_obj": c["obj"],
        "meta_frames": np.array(c["frames"]), "meta_N": c["N"], "meta_M": c["M"], "meta_L": c["L"],
        "meta_hp_keys": np.array(sorted(c["hp"].keys())), "meta_hp_vals": np.array([c["hp"][k] for k in sorted(c["hp"].keys())]),
        "matches0": pred["matches0"].numpy(), "matches1": pred["matches1"].numpy(),
        "matching_scores0": pred["matching_scores0"].numpy(), "matching_scores1": pred["matching_scores1"].numpy(),
        "raw_indices0": conf.argmax(2), "raw_indices1": conf.argmax(1),
        "conf_rowmax": conf.max(2), "conf_colmax": conf.max(1),
        "conf_rowsum": conf.astype(np.float64).sum(2), "conf_colsum": conf.astype(np.float64).sum(1),
        "conf64_rowmax": conf64.numpy().max(2),
        "max_abs_conf32_vs_conf64": np.abs(conf.astype(np.float64) - conf64.numpy()).max(),
    }
    if c["full"]:
        out["conf_matrix"] = conf
    else:
        out["conf_sample_rows"] = np.arange(0, c["N"], 7)
        out["conf_sample_cols"] = np.arange(0, c["M"], 5)
        out["conf_sample"] = conf[:, ::7, ::5]
    nm = int((out["matches0"] > -1).sum())
    print(f"{c['name']}: matches0 valid={nm}/{c['N']} conf.max={conf.max():.4f} "
          f"|conf32-conf64|max={out['max_abs_conf32_vs_conf64']:.2e} dtypes "
          f"{pred['matches0'].dtype} {pred['matching_scores0'].dtype} {tuple(conf.shape)}")
    np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), **out)


def empty_case():
    hp = dict(synthetic.DEFAULT_HPARAMS)
    model = GATsSuperGlue(hp).eval()
    data = synthetic.make_batch(1, [11], 0, 96, 8)
    with torch.no_grad():
        ret = model({k: torch.from_numpy(v) for k, v in data.items()})
    assert isinstance(ret, dict)
    np.savez_compressed(os.path.join(HERE, "empty_n0_m96.npz"),
                        keys=np.array(sorted(ret.keys())),
                        matches0=ret["matches0"].numpy(), matches1=ret["matches1"].numpy(),
                        matching_scores0=ret["matching_scores0"].numpy(),
                        matching_scores1=ret["matching_scores1"].numpy(),
                        skip_train=np.array(ret["skip_train"]))
    print("empty:", {k: (tuple(v.shape), v.dtype) if hasattr(v, "shape") else v for k, v in ret.items()})


def _reference_functions(path, names):
    """Source text of top-level functions of a reference module that cannot be imported here (feature_process.py needs
    h5py at import time), exec'd with numpy in scope: the functions run UNMODIFIED."""
    import ast
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return [ns[n] for n in names]


def mean_cases():
    """mean_descriptors / mean_scores run from the reference's own source text (feature_process.py:297-317)."""
    mean_descriptors, mean_scores = _reference_functions(
        "/root/reference/src/sfm/postprocess/feature_process.py", ["mean_descriptors", "mean_scores"])
    for name, seed, M, max_len in [("mean_descriptors_m300", 7, 300, 12), ("mean_tracks_long_m48", 11, 48, 400)]:
        desc, idxs = synthetic.make_tracks(seed, M, max_len=max_len)
        scores = synthetic.make_track_scores(seed, idxs)
        avg = mean_descriptors(desc, idxs)
        avg_s = mean_scores(scores, idxs)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, M=M, max_len=max_len, avg=avg, avg_scores=avg_s)
        print(name, avg.shape, avg.dtype, avg_s.shape, int(idxs.max()))


def features3d_cases():
    """pad_features3d_random / build_features3d_leaves of the imported reference (src/utils/data_utils.py:143-205) under a fixed
    np.random.seed, and the reference forward on an object built by them (all-ones dustbin leaves, duplicate all-ones padded 3D
    points -> exactly tied confidences)."""
    from src.utils import data_utils
    n_points, L = 40, 8
    obs, obs_scores, idxs, avg, avg_scores = synthetic.make_sfm_features(21, n_points)
    out = {"seed": 21, "n_points": n_points, "num_leaf": L, "np_seed": 123}
    for tag, n_target in [("same", n_points), ("pad", n_points + 8), ("trunc", n_points - 6)]:
        d3, s3 = data_utils.pad_features3d_random(avg, avg_scores, n_target)
        np.random.seed(123)
        d2, s2 = data_utils.build_features3d_leaves(obs, obs_scores, idxs, n_target, L)
        out[f"{tag}_n_target"] = n_target
        out[f"{tag}_desc3d"], out[f"{tag}_scores3d"] = d3.numpy(), s3.numpy()
        out[f"{tag}_desc2d"], out[f"{tag}_scores2d"] = d2.numpy(), s2.numpy()
        print("features3d", tag, tuple(d3.shape), tuple(d2.shape), d3.dtype)
    np.savez_compressed(os.path.join(HERE, "features3d_n40_l8.npz"), **out)

    # forward on a padded object: 88 real points padded to 96 (8 identical all-ones columns), dustbin leaves
    n_real, n_target, N = 88, 96, 64
    obs, obs_scores, idxs, avg, avg_scores = synthetic.make_sfm_features(33, n_real, max_len=11)
    d3, _ = data_utils.pad_features3d_random(avg, avg_scores, n_target)
    np.random.seed(5)
    d2, _ = data_utils.build_features3d_leaves(obs, obs_scores, idxs, n_target, L)
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0, damped=True, hparams=hp)
    model = GATsSuperGlue(hp).eval()
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    q, _ = synthetic.make_frame(77, avg, N)
    inp = {"keypoints2d": torch.zeros(1, N, 2), "keypoints3d": torch.zeros(1, n_target, 3),
           "descriptors2d_query": torch.from_numpy(q)[None], "descriptors3d_db": d3[None], "descriptors2d_db": d2[None]}
    with torch.no_grad():
        pred, conf = model(inp)
    conf = conf.numpy()
    np.savez_compressed(os.path.join(HERE, "dustbin_n64_m96.npz"), desc3d=d3.numpy(), desc2d=d2.numpy(), query=q, n_real=n_real,
                        matches0=pred["matches0"].numpy(), matches1=pred["matches1"].numpy(),
                        matching_scores0=pred["matching_scores0"].numpy(), matching_scores1=pred["matching_scores1"].numpy(),
                        conf_matrix=conf, raw_indices0=conf.argmax(2), raw_indices1=conf.argmax(1))
    print("dustbin forward: matches", int((pred["matches0"] > -1).sum()), "conf.max", float(conf.max()),
          "tied padded columns identical:", bool((conf[0][:, n_real:] == conf[0][:, n_real:n_real + 1]).all()))


PNP_SCENES = [  # seed, n correspondences, outlier fraction
    (1, 300, 0.3), (2, 500, 0.5), (3, 120, 0.6), (4, 60, 0.2), (5, 800, 0.4), (6, 30, 0.5), (7, 6, 0.0), (8, 3, 0.0)]


def pnp_cases():
    """ransac_PnP of the imported, unmodified reference (src/utils/eval_utils.py:18-42; cv2 of THIS image) on seeded scenes."""
    from src.utils import eval_utils
    import cv2
    out = {"cv2_version": cv2.__version__, "scenes": np.array(PNP_SCENES)}
    for seed, n, frac in PNP_SCENES:
        K, uv, P, gt = synthetic.make_pnp_scene(seed, n, frac)
        pose, pose_homo, inliers = eval_utils.ransac_PnP(K, uv, P.copy(), scale=1000)
        out[f"pose_{seed}"] = pose
        out[f"n_inliers_{seed}"] = len(inliers)
        out[f"gt_{seed}"] = gt
        r = np.rad2deg(np.arccos(np.clip((np.trace(pose[:, :3] @ gt[:, :3].T) - 1) / 2, -1, 1)))
        print(f"pnp scene {seed}: n={n} outliers={frac}: reference inliers {len(inliers)}, err vs gt {r:.3f} deg {np.linalg.norm(pose[:, 3] - gt[:, 3]) * 100:.3f} cm")
    np.savez_compressed(os.path.join(HERE, "pnp_scenes.npz"), **out)


def state_dict_spec():
    """Key names / shapes of the reference module's state dict (load_state_dict compatibility)."""
    import json
    m = GATsSuperGlue(dict(synthetic.DEFAULT_HPARAMS))
    spec = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f, indent=0)
    print("state_dict_spec:", len(spec), "keys")


SUPERPOINT_CASES = [
    # name, weight seed, score gain, image seeds, H, W, config, descriptor columns stored (every n-th key point)
    dict(name="superpoint_b2_128x160", wseed=0, gain=4.0, images=[1, 2], H=128, W=160, conf=dict(synthetic.SUPERPOINT_CONF), every=1),
    dict(name="superpoint_topk_96x128", wseed=1, gain=4.0, images=[3], H=96, W=128,
         conf={"nms_radius": 4, "max_keypoints": 150, "keypoint_threshold": 0.005, "remove_borders": 4}, every=1),
    dict(name="superpoint_all_64x64", wseed=2, gain=1.0, images=[4], H=64, W=64, conf={"max_keypoints": -1, "remove_borders": 2}, every=1),
    dict(name="superpoint_512x512", wseed=0, gain=4.0, images=[5], H=512, W=512, conf=dict(synthetic.SUPERPOINT_CONF), every=16),
]


def superpoint_cases():
    """SuperPoint.forward of the imported, unmodified reference module on seeded weights and images."""
    import json
    import warnings
    from src.models.extractors.SuperPoint.superpoint import SuperPoint
    with open(os.path.join(HERE, "superpoint_state_dict_spec.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in SuperPoint({}).state_dict().items()}, f, indent=0)
    for c in SUPERPOINT_CASES:
        sd = synthetic.make_superpoint_state_dict(c["wseed"], c["gain"])
        model = SuperPoint(c["conf"]).eval()
        model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
        img = torch.from_numpy(np.stack([synthetic.make_image(i, c["H"], c["W"]) for i in c["images"]], 0))
        out = {"meta_wseed": c["wseed"], "meta_gain": c["gain"], "meta_images": np.array(c["images"]), "meta_H": c["H"], "meta_W": c["W"],
               "meta_conf_keys": np.array(sorted(c["conf"].keys())), "meta_conf_vals": np.array([float(c["conf"][k]) for k in sorted(c["conf"].keys())]),
               "meta_every": c["every"]}
        real_version = torch.__version__
        for tag, version in (("ac0", real_version), ("ac1", "1.8.0")):
            torch.__version__ = version
            try:
                with torch.no_grad(), warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    pred = model(img)
            finally:
                torch.__version__ = real_version
            for b in range(len(c["images"])):
                if tag == "ac0":
                    out[f"keypoints_{b}"] = pred["keypoints"][b].numpy()
                    out[f"scores_{b}"] = pred["scores"][b].numpy()
                else:
                    assert np.array_equal(out[f"keypoints_{b}"], pred["keypoints"][b].numpy())
                out[f"descriptors_{tag}_{b}"] = pred["descriptors"][b].numpy()[:, ::c["every"]]
        n = [len(out[f"keypoints_{b}"]) for b in range(len(c["images"]))]
        d = float(np.abs(out["descriptors_ac0_0"] - out["descriptors_ac1_0"]).max())
        print(f"{c['name']}: key points {n}, score range [{out['scores_0'].min():.4f}, {out['scores_0'].max():.4f}], "
              f"|desc(ac=False) - desc(ac=True)| max {d:.3f}")
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    what = set(sys.argv[1:]) or {"matcher", "producers", "pnp", "superpoint"}
    if "matcher" in what:
        state_dict_spec()
        for c in CASES:
            run_case(c)
        empty_case()
    if "producers" in what:
        mean_cases()
        features3d_cases()
    if "pnp" in what:
        pnp_cases()
    if "superpoint" in what:
        superpoint_cases()
