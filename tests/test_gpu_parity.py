"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the module / C ABI
against (a) the golden vectors produced by the reference module and (b) the CPU oracle on
the same seeded inputs.  Contract (BASELINE.json north_star): conf within 1e-4 abs,
match indices bit-exact."""
import os

import numpy as np
import pytest
import torch

from onepose_b200 import GATsSuperGlue, synthetic
from oracle import gats_spg_oracle as oracle
from tests.golden_util import RELEASED_CASES, conf_reference_view, load_case

pytestmark = pytest.mark.gpu

CONF_TOL = 1e-4          # north_star tolerance
BACKENDS = ["tcgen05", "simt", "tcgen05_unfused"]


def _module(sd, hp, backend):
    m = GATsSuperGlue(dict(hp), gemm_backend=backend).eval()
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    return m.cuda()


def _cuda(data):
    return {k: torch.from_numpy(v).cuda() for k, v in data.items()}


def _check_against(out, ref, tag):
    """out: batched dict from the CUDA path (torch, cuda); ref: oracle dict (torch cpu)."""
    conf = out["conf_matrix"].cpu()
    rc = ref["conf_matrix"]
    err = float((conf - rc).abs().max())
    assert err <= CONF_TOL, f"{tag}: max|dconf| = {err:.3e}"
    np.testing.assert_array_equal(out["matches0"].cpu().numpy(), ref["matches0"].numpy(), err_msg=tag)
    np.testing.assert_array_equal(out["matches1"].cpu().numpy(), ref["matches1"].numpy(), err_msg=tag)
    np.testing.assert_allclose(out["matching_scores0"].cpu().numpy(), ref["matching_scores0"].numpy(), atol=CONF_TOL)
    np.testing.assert_allclose(out["matching_scores1"].cpu().numpy(), ref["matching_scores1"].numpy(), atol=CONF_TOL)
    # raw arg-max (before mutual/threshold): must agree wherever the oracle's decision is not a
    # floating-point coin flip (top-1 vs top-2 separated by more than 1e-5 relative)
    top2 = rc.topk(2, dim=2).values
    decided = (top2[..., 0] - top2[..., 1]) > 1e-5 * top2[..., 0]
    mine = conf.argmax(2)
    assert bool((mine == ref["raw_indices0"])[decided].all()), tag
    top2c = rc.topk(2, dim=1).values
    decided_c = (top2c[:, 0] - top2c[:, 1]) > 1e-5 * top2c[:, 0]
    assert bool((conf.argmax(1) == ref["raw_indices1"])[decided_c].all()), tag
    return err


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", RELEASED_CASES)
def test_golden_reference_outputs(name, backend):
    g, hp, sd, data = load_case(name)
    m = _module(sd, hp, backend)
    pred, conf = m(_cuda(data))
    ref_conf, mine = conf_reference_view(g, conf.cpu().numpy())
    assert np.abs(ref_conf - mine).max() <= CONF_TOL
    assert pred["matches0"].dtype == torch.int64 and pred["matching_scores0"].dtype == torch.float32
    np.testing.assert_array_equal(pred["matches0"].cpu().numpy(), g["matches0"])      # bit-exact
    np.testing.assert_array_equal(pred["matches1"].cpu().numpy(), g["matches1"])
    np.testing.assert_allclose(pred["matching_scores0"].cpu().numpy(), g["matching_scores0"], atol=CONF_TOL)
    np.testing.assert_allclose(pred["matching_scores1"].cpu().numpy(), g["matching_scores1"], atol=CONF_TOL)
    assert tuple(conf.shape) == (len(g["meta_frames"]), int(g["meta_N"]), int(g["meta_M"]))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("hp_over", [{"include_self": False}, {"additional": True}])
def test_gats_variants_golden(hp_over, backend):
    name = "noself_n64_m96" if "include_self" in hp_over else "additional_n64_m96"
    g, hp, sd, data = load_case(name)
    m = _module(sd, hp, backend)
    pred, conf = m(_cuda(data))
    assert np.abs(g["conf_matrix"] - conf.cpu().numpy()).max() <= CONF_TOL
    np.testing.assert_array_equal(pred["matches0"].cpu().numpy(), g["matches0"])


def test_with_linear_transform_is_refused_loudly():
    hp = dict(synthetic.DEFAULT_HPARAMS, with_linear_transform=True)
    m = GATsSuperGlue(hp).cuda()
    with pytest.raises(Exception, match="with_linear_transform"):
        m(_cuda(synthetic.make_batch(1, [1], 32, 64, 8)))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("B,N,M,L,damped", [(3, 300, 700, 8, True), (1, 129, 257, 5, True), (2, 17, 40, 8, False),
                                            (5, 256, 384, 8, True)])
def test_oracle_parity_batched(B, N, M, L, damped, backend):
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(4, damped=damped)
    data = synthetic.make_batch(9, list(range(100, 100 + B)), N, M, L)
    ref = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    m = _module(sd, hp, backend)
    m.set_chunk_frames(2)            # exercise the chunk loop (3 = 2 + 1)
    m(_cuda(data))
    _check_against(m.last_batched, ref, f"B{B} N{N} M{M} L{L}")


@pytest.mark.parametrize("backend", BACKENDS)
def test_different_objects_in_one_batch(backend):
    """The reference forward accepts per-element 3D descriptors; the drop-in groups by object."""
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    d1 = synthetic.make_batch(1, [5, 6], 96, 160, 8)
    d2 = synthetic.make_batch(2, [7], 96, 160, 8)
    data = {k: np.concatenate([d1[k], d2[k]], 0) for k in d1}
    ref = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    m = _module(sd, hp, backend)
    pred, conf = m(_cuda(data))
    _check_against(m.last_batched, ref, "mixed objects")


def test_metric_shape_one_frame_vs_oracle():
    """BASELINE metric shape N2D=1024, N3D=7000 (one frame: the oracle needs ~1-2 s of CPU)."""
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    data = synthetic.make_batch(7, [70], 1024, 7000, 8)
    ref = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    m = _module(sd, hp, "tcgen05")
    m(_cuda(data))
    err = _check_against(m.last_batched, ref, "metric shape")
    assert int((m.last_batched["matches0"] > -1).sum()) >= 400      # planted correspondences recovered
    print(f"metric-shape max|dconf| = {err:.2e}")


def test_dense_stress_shape_vs_oracle():
    """BASELINE configs[3]: N2D=2000, N3D=15000 (dual-softmax + mutual-NN at the largest size the survey names)."""
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    data = synthetic.make_batch(11, [90], 2000, 15000, 8)
    ref = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    m = _module(sd, hp, "tcgen05")
    m(_cuda(data))
    err = _check_against(m.last_batched, ref, "dense stress")
    assert int((m.last_batched["matches0"] > -1).sum()) == int((ref["matches0"] > -1).sum()) > 500
    print(f"dense-stress max|dconf| = {err:.2e}")


def test_full_size_properties():
    """Size-independent properties at the bench configuration (B=8 frames of one object)."""
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    B, N, M = 8, 1024, 7000
    data = synthetic.make_batch(7, list(range(70, 70 + B)), N, M, 8)
    m = _module(sd, hp, "tcgen05")
    m(_cuda(data))
    out = m.last_batched
    conf = out["conf_matrix"]
    # softmax factors: row sums of conf <= max col-softmax <= 1, every entry in [0, 1]
    assert float(conf.min()) >= 0.0 and float(conf.max()) <= 1.0 + 1e-6
    assert float(conf.sum(2).max()) <= 1.0 + 1e-4 and float(conf.sum(1).max()) <= 1.0 + 1e-4
    m0, m1 = out["matches0"], out["matches1"]
    ar = torch.arange(N, device=conf.device)[None].expand(B, N)
    valid = m0 > -1
    # mutual consistency: matches1[matches0[n]] == n for valid n; scores agree with conf
    assert bool((m1.gather(1, m0.clamp_min(0))[valid] == ar[valid]).all())
    picked = conf.gather(2, m0.clamp_min(0)[..., None])[..., 0]
    assert torch.equal(picked[valid], out["matching_scores0"][valid])
    assert bool((out["matching_scores0"][valid] > hp["match_threshold"]).all())
    # frame b of a batch == the same frame run alone (frames are independent)
    solo = _module(sd, hp, "tcgen05")
    one = {k: v[3:4] for k, v in data.items()}
    solo(_cuda(one))
    assert torch.equal(solo.last_batched["matches0"][0], m0[3])
    assert float((solo.last_batched["conf_matrix"][0] - conf[3]).abs().max()) <= 1e-6
    # planted correspondences (first N/2 queries) are found
    for b in range(B):
        _, perm = synthetic.make_frame(70 + b, data["descriptors3d_db"][0], N)
        got = m0[b, : N // 2].cpu().numpy()
        assert (got == perm).mean() > 0.95


def test_gemm_cores_agree_and_match_fp64():
    """tcgen05 3-pass fp16-split GEMM vs the SIMT fp32 cross-check vs numpy fp64, same split operands."""
    import ctypes as C
    from onepose_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(0)
    for rows, n_out, K in [(128, 256, 64), (256, 256, 256), (384, 768, 256), (256, 512, 512), (1024, 7168, 256)]:
        a = torch.from_numpy((rs.randn(rows, K) * np.exp(rs.randn(rows, 1))).astype(np.float32)).cuda()
        b = torch.from_numpy((rs.randn(n_out, K) / np.sqrt(K)).astype(np.float32)).cuda()
        planes = [torch.empty_like(a, dtype=torch.float16) for _ in range(2)] + [torch.empty_like(b, dtype=torch.float16) for _ in range(2)]
        assert lib.opb_debug_split(a.data_ptr(), planes[0].data_ptr(), planes[1].data_ptr(), a.numel(), None) == 0
        assert lib.opb_debug_split(b.data_ptr(), planes[2].data_ptr(), planes[3].data_ptr(), b.numel(), None) == 0
        outs = []
        for backend in (0, 1):
            c = torch.zeros(rows, n_out, dtype=torch.float32, device="cuda")
            rc = lib.opb_debug_gemm(*(p.data_ptr() for p in planes), c.data_ptr(), rows, n_out, K, backend, None)
            assert rc == 0, (backend, rc)
            torch.cuda.synchronize()
            outs.append(c.cpu().double())
        ah = (planes[0].cpu().double() + planes[1].cpu().double()) / 64.0
        bh = (planes[2].cpu().double() + planes[3].cpu().double()) / 64.0
        ref = ah @ bh.T
        scale = float(ref.abs().max())
        e_tc = float((outs[0] - ref).abs().max()) / scale
        e_simt = float((outs[1] - ref).abs().max()) / scale
        print(f"gemm {rows}x{n_out}x{K}: rel err tcgen05 {e_tc:.2e}, simt {e_simt:.2e}")
        # the tensor core truncates (does not round) its fp32 accumulator: ~2^-25 relative per accumulation step
        assert e_simt < 2e-6 and e_tc < 2e-6 + 1.2e-8 * K
        # the split itself represents fp32 to ~2^-22
        assert float((ah.float() - a.cpu()).abs().max() / a.abs().max()) < 2.0 ** -21


def test_segmented_mean_matches_reference_golden():
    import ctypes as C
    from onepose_b200 import _lib
    from tests.golden_util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "mean_descriptors_m300.npz"))
    desc, idxs = synthetic.make_tracks(int(g["seed"]), int(g["M"]))
    d = torch.from_numpy(desc).cuda()
    l = torch.from_numpy(idxs).cuda()
    out = torch.empty(len(idxs), desc.shape[1], dtype=torch.float64, device="cuda")
    rc = _lib.load().opb_segmented_mean_f64(d.data_ptr(), l.data_ptr(), len(idxs), desc.shape[1], out.data_ptr(), None)
    assert rc == 0
    np.testing.assert_allclose(out.cpu().numpy(), g["avg"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("backend", BACKENDS)
def test_object_prologue_hoisting_matches_per_frame_evaluation(backend):
    """Layers 0-1 of the 3D side are frame-invariant: evaluating them once per call (default) must give the
    same answer as evaluating them per frame like the reference (GATs_SuperGlue.py:50-64)."""
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    data = synthetic.make_batch(4, [1, 2, 3], 200, 500, 8)
    ref = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    outs = []
    for hoist in (True, False):
        m = _module(sd, hp, backend)
        m.set_hoist(hoist)
        m.set_chunk_frames(2)
        m(_cuda(data))
        _check_against(m.last_batched, ref, f"hoist={hoist}")
        outs.append(m.last_batched)
    assert float((outs[0]["conf_matrix"] - outs[1]["conf_matrix"]).abs().max()) <= 2e-6
    assert torch.equal(outs[0]["matches0"], outs[1]["matches0"])


@pytest.mark.parametrize("level", [0, 1, 2])
def test_fuse_levels_agree(level):
    """Every epilogue-fusion level of the tcgen05 path gives the oracle's answer."""
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    data = synthetic.make_batch(6, [3, 4, 5], 333, 700, 8)
    ref = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    m = _module(sd, hp, "tcgen05")
    m.set_fuse_level(level)
    m(_cuda(data))
    _check_against(m.last_batched, ref, f"fuse level {level}")


def test_host_buffer_call_equals_device_call():
    """opb_forward_host (pinned host in / host out, chunked H2D on a side stream) == opb_forward on device tensors."""
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    data = synthetic.make_batch(8, list(range(5)), 260, 600, 8)
    m = _module(sd, hp, "tcgen05")
    m.set_chunk_frames(2)                       # 3 chunks: 2 + 2 + 1
    m.set_object(torch.from_numpy(data["descriptors3d_db"][0]).cuda(), torch.from_numpy(data["descriptors2d_db"][0]).cuda())
    q = torch.from_numpy(data["descriptors2d_query"])
    dev = m.match_frames(q.cuda())
    for _ in range(2):                          # second call reuses the staging buffers / events
        host = m.match_frames_host(q.pin_memory())
    for k in ("matches0", "matches1", "matching_scores0", "matching_scores1"):
        assert torch.equal(host[k], dev[k].cpu()), k
    ref = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    np.testing.assert_array_equal(host["matches0"].numpy(), ref["matches0"].numpy())


def test_operand_range_guard_fires():
    """Activations beyond the fp16-split operand range (|x| >= 1023) must be reported, not silently wrong."""
    from onepose_b200 import _lib
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    sd = {k: (v * 3e4 if k.endswith("mlp.3.weight") else v) for k, v in sd.items()}     # residual stream explodes
    m = _module(sd, hp, "tcgen05")
    with pytest.raises(_lib.OpbError, match="OPB_E_RANGE"):
        m(_cuda(synthetic.make_batch(1, [1], 128, 256, 8)))
    # the guard is cleared by the check: a sane model on the same process still works
    ok = _module(synthetic.make_state_dict(0), hp, "tcgen05")
    pred, _ = ok(_cuda(synthetic.make_batch(1, [1], 128, 256, 8)))
    assert int((pred["matches0"] > -1).sum()) > 0


def test_repeat_calls_are_deterministic():
    hp = dict(synthetic.DEFAULT_HPARAMS)
    sd = synthetic.make_state_dict(0)
    data = _cuda(synthetic.make_batch(3, [1, 2], 200, 500, 8))
    m = _module(sd, hp, "tcgen05")
    m(data)
    a = {k: v.clone() for k, v in m.last_batched.items()}
    m(data)
    for k, v in m.last_batched.items():
        assert torch.equal(a[k], v), k
