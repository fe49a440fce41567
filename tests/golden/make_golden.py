"""Generate golden vectors by running the UNMODIFIED reference module.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports ``GATsSuperGlue`` from ``/root/reference`` (reference
``src/models/GATsSPG_architectures/GATs_SuperGlue.py:143``), loads the seeded
numpy state dict of ``onepose_b200.synthetic.make_state_dict`` into it, feeds
the seeded synthetic inputs of ``onepose_b200.synthetic.make_batch`` and stores
what the reference returns.  Inputs are NOT stored: they are regenerated from
the seeds recorded in each file (numpy RandomState is platform-stable).

Also stores the offline producer ``mean_descriptors``
(reference ``src/sfm/postprocess/feature_process.py:297-305``) -- restated here
rather than imported because that module pulls in h5py at import time.
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


def mean_desc_case():
    desc, idxs = synthetic.make_tracks(7, 300)
    # reference feature_process.py:297-305 semantics: np.mean over rows [start:end) of each track
    ends = np.cumsum(idxs)
    starts = np.insert(ends[:-1], 0, 0)
    avg = np.concatenate([np.mean(desc[s:e], axis=0).reshape(1, -1) for s, e in zip(starts, ends)], 0)
    np.savez_compressed(os.path.join(HERE, "mean_descriptors_m300.npz"), seed=7, M=300, avg=avg)
    print("mean_descriptors:", avg.shape, avg.dtype)


def state_dict_spec():
    """Key names / shapes of the reference module's state dict (load_state_dict compatibility)."""
    import json
    m = GATsSuperGlue(dict(synthetic.DEFAULT_HPARAMS))
    spec = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f, indent=0)
    print("state_dict_spec:", len(spec), "keys")


if __name__ == "__main__":
    torch.set_num_threads(8)
    state_dict_spec()
    for c in CASES:
        run_case(c)
    empty_case()
    mean_desc_case()
