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


def test_oracle_mean_descriptors():
    g = np.load(f"{GOLDEN_DIR}/mean_descriptors_m300.npz")
    desc, idxs = synthetic.make_tracks(int(g["seed"]), int(g["M"]))
    np.testing.assert_allclose(oracle.mean_descriptors(desc, idxs), g["avg"], rtol=0, atol=1e-15)
