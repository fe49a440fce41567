"""Pin the oracle (oracle/gats_spg_oracle.py) to the reference's own outputs
(tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import gats_spg_oracle as oracle
from onepose_b200 import synthetic
from tests.golden_util import FORWARD_CASES, GOLDEN_DIR, conf_reference_view, load_case


@pytest.mark.parametrize("name", FORWARD_CASES)
def test_oracle_fp32_matches_reference(name):
    g, hp, sd, data = load_case(name)
    out = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    conf = out["conf_matrix"].numpy()
    ref, mine = conf_reference_view(g, conf)
    # same fp32 op sequence up to BLAS blocking: far inside the 1e-4 contract
    assert np.abs(ref - mine).max() < 2e-5
    # reference pred = batch element 0 (GATs_SuperGlue.py:232-237); dtypes int64 / fp32
    assert out["matches0"].dtype == torch.int64
    np.testing.assert_array_equal(out["matches0"][0].numpy(), g["matches0"])
    np.testing.assert_array_equal(out["matches1"][0].numpy(), g["matches1"])
    np.testing.assert_allclose(out["matching_scores0"][0].numpy(), g["matching_scores0"], atol=2e-5)
    np.testing.assert_allclose(out["matching_scores1"][0].numpy(), g["matching_scores1"], atol=2e-5)
    np.testing.assert_array_equal(out["raw_indices0"].numpy(), g["raw_indices0"])
    np.testing.assert_array_equal(out["raw_indices1"].numpy(), g["raw_indices1"])
    np.testing.assert_allclose(conf.astype(np.float64).sum(2), g["conf_rowsum"], atol=1e-4)
    np.testing.assert_allclose(conf.astype(np.float64).sum(1), g["conf_colsum"], atol=1e-4)


@pytest.mark.parametrize("name", ["tiny_n64_m96", "ragged_b2_n200_m333"])
def test_oracle_fp64_is_referee(name):
    g, hp, sd, data = load_case(name)
    out = oracle.forward(oracle.params_from_numpy(sd, torch.float64), data, hp, dtype=torch.float64)
    np.testing.assert_allclose(out["conf_matrix"].numpy().max(2), g["conf64_rowmax"], atol=1e-9)


def test_oracle_empty_input_returns_bare_dict():
    g = np.load(f"{GOLDEN_DIR}/empty_n0_m96.npz")
    data = synthetic.make_batch(1, [11], 0, 96, 8)
    sd = synthetic.make_state_dict(0)
    out = oracle.forward(oracle.params_from_numpy(sd), data, synthetic.DEFAULT_HPARAMS)
    assert isinstance(out, dict) and sorted(out.keys()) == list(g["keys"])
    assert out["matches0"].dtype == torch.int32 and out["matches0"].shape[1:] == g["matches0"].shape
    assert out["skip_train"] is True


def test_oracle_rejects_other_match_types():
    hp = dict(synthetic.DEFAULT_HPARAMS, match_type="sinkhorn")
    with pytest.raises(NotImplementedError):
        oracle.forward({}, synthetic.make_batch(1, [1], 8, 8, 2), hp)


@pytest.mark.parametrize("name", ["mean_descriptors_m300", "mean_tracks_long_m48"])
def test_oracle_segmented_means_match_reference_source(name):
    """mean_descriptors / mean_scores: the goldens are outputs of the reference's own source text (make_golden.mean_cases)."""
    g = np.load(f"{GOLDEN_DIR}/{name}.npz")
    desc, idxs = synthetic.make_tracks(int(g["seed"]), int(g["M"]), max_len=int(g["max_len"]))
    scores = synthetic.make_track_scores(int(g["seed"]), idxs)
    np.testing.assert_array_equal(oracle.mean_descriptors(desc, idxs), g["avg"])
    np.testing.assert_array_equal(oracle.mean_scores(scores, idxs), g["avg_scores"])


def test_oracle_features3d_match_reference():
    """pad_features3d_random / build_features3d_leaves: byte-identical to the imported reference under the same np.random.seed."""
    g = np.load(f"{GOLDEN_DIR}/features3d_n40_l8.npz")
    obs, obs_scores, idxs, avg, avg_scores = synthetic.make_sfm_features(int(g["seed"]), int(g["n_points"]))
    for tag in ("same", "pad", "trunc"):
        nt = int(g[f"{tag}_n_target"])
        d3, s3 = oracle.pad_features3d_random(avg, avg_scores, nt)
        np.random.seed(int(g["np_seed"]))
        d2, s2 = oracle.build_features3d_leaves(obs, obs_scores, idxs, nt, int(g["num_leaf"]))
        for mine, key in ((d3, "desc3d"), (s3, "scores3d"), (d2, "desc2d"), (s2, "scores2d")):
            assert mine.dtype == np.float32 and mine.tobytes() == g[f"{tag}_{key}"].tobytes(), (tag, key)


def test_oracle_dustbin_object_matches_reference():
    """Reference-built object with all-ones dustbin leaves and duplicate padded 3D points: exactly tied confidences."""
    from tests.golden_util import load_dustbin_case
    g, hp, sd, data = load_dustbin_case()
    out = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    assert np.abs(out["conf_matrix"].numpy() - g["conf_matrix"]).max() < 2e-5
    np.testing.assert_array_equal(out["matches0"][0].numpy(), g["matches0"])
    np.testing.assert_array_equal(out["matches1"][0].numpy(), g["matches1"])
    n_real = int(g["n_real"])
    assert (g["conf_matrix"][0][:, n_real:] == g["conf_matrix"][0][:, n_real:n_real + 1]).all()     # the tie is exact in the reference


def test_oracle_pnp_matches_reference_function():
    """ransac_PnP restated (oracle/pnp_oracle.py) vs the imported reference function's outputs (same cv2, fixed-seed RNG)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import pnp_oracle
    g = np.load(f"{GOLDEN_DIR}/pnp_scenes.npz")
    if str(g["cv2_version"]) != cv2.__version__:
        pytest.skip("golden made with another OpenCV build")
    for seed, n, frac in g["scenes"]:
        K, uv, P, gt = synthetic.make_pnp_scene(int(seed), int(n), float(frac))
        pose, homo, inliers = pnp_oracle.ransac_PnP(K, uv, P, scale=1000)
        np.testing.assert_allclose(pose, g[f"pose_{int(seed)}"], atol=1e-12)
        assert len(inliers) == int(g[f"n_inliers_{int(seed)}"])
        assert homo.shape == (4, 4)
