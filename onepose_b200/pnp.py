"""Object pose from the matcher's 2D-3D correspondences -- GPU RANSAC-PnP.

``ransac_PnP`` mirrors the reference's function of the same name (``src/utils/eval_utils.py:18-42``: same arguments, same
return triple ``(pose [3,4], pose_homo [4,4], inliers)`` as numpy, identity pose and ``[]`` when no pose can be estimated);
``ransac_pnp_batch`` solves many frames in one launch and keeps everything on the device.  The arithmetic is
``csrc/pnp_ransac.cu`` (P3P minimal samples, inlier counting over all correspondences, Gauss-Newton refinement on the
inliers) behind ``opb_ransac_pnp``.  The reference delegates to ``cv2.solvePnPRansac`` (EPnP, 10000 iterations, 5 px);
parity is on the pose (the reference's cm-degree metric), not on bits -- OpenCV draws its own random samples.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

DEFAULT_HYPOTHESES = 4096      # minimal samples per frame: P(no all-inlier sample) < 1e-9 down to 15 % inliers


def ransac_pnp_batch(K: torch.Tensor, pts2d: torch.Tensor, pts3d: torch.Tensor, offsets: torch.Tensor, reproj_error: float = 5.0,
                     hypotheses: int = DEFAULT_HYPOTHESES, seed: int = 0):
    """K [B,3,3], pts2d [total,2], pts3d [total,3], offsets [B+1] (frame b owns [offsets[b], offsets[b+1])) -- CUDA tensors
    (any float dtype; computed in fp64).  Returns (pose [B,3,4] f64, inlier_mask [total] bool, n_inliers [B] int32), on the device."""
    lib = _lib.load()
    dev = pts2d.device
    if dev.type != "cuda":
        raise RuntimeError("onepose_b200 has no CPU path: tensors must be on a CUDA device")
    B = int(offsets.numel()) - 1
    K = K.to(device=dev, dtype=torch.float64).reshape(B, 9).contiguous()
    p2 = pts2d.to(torch.float64).contiguous()
    p3 = pts3d.to(torch.float64).contiguous()
    off = offsets.to(device=dev, dtype=torch.int32).contiguous()
    total = int(p2.shape[0])
    ws = torch.empty(B * hypotheses * 13, dtype=torch.float64, device=dev)
    pose = torch.empty(B, 12, dtype=torch.float64, device=dev)
    mask = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(lib.opb_ransac_pnp(K.data_ptr(), p2.data_ptr(), p3.data_ptr(), off.data_ptr(), B, int(hypotheses), float(reproj_error),
                                      int(seed) & 0xFFFFFFFFFFFFFFFF, ws.data_ptr(), pose.data_ptr(), mask.data_ptr(), cnt.data_ptr(), st))
    return pose.reshape(B, 3, 4), mask[:total].bool(), cnt


def ransac_PnP(K, pts_2d, pts_3d, scale=1, device=None):
    """Drop-in for the reference ``ransac_PnP`` (eval_utils.py:18-42): numpy in, numpy out.
    ``scale`` only conditions OpenCV's EPnP in the reference (points x scale, translation / scale): applied and undone here
    too so that the argument keeps its meaning.  Returns (pose [3,4], pose_homo [4,4], inliers [k,1] int32 or [])."""
    dev = torch.device(device if device is not None else "cuda")
    pts_2d = np.ascontiguousarray(np.asarray(pts_2d, dtype=np.float64)).reshape(-1, 2)
    pts_3d = np.ascontiguousarray(np.asarray(pts_3d, dtype=np.float64)).reshape(-1, 3) * scale
    n = pts_2d.shape[0]
    if n < 4:                                                   # cv2.error branch of the reference (:40-42)
        return np.eye(4)[:3], np.eye(4), []
    pose, mask, cnt = ransac_pnp_batch(torch.from_numpy(np.asarray(K, dtype=np.float64))[None].to(dev), torch.from_numpy(pts_2d).to(dev),
                                       torch.from_numpy(pts_3d).to(dev), torch.tensor([0, n], dtype=torch.int32, device=dev))
    if int(cnt[0]) == 0:
        return np.eye(4)[:3], np.eye(4), []
    pose = pose[0].cpu().numpy()
    pose[:, 3] /= scale
    pose_homo = np.concatenate([pose, np.array([[0, 0, 0, 1.0]])], axis=0)
    inliers = torch.nonzero(mask).to(torch.int32).cpu().numpy().reshape(-1, 1)
    return pose, pose_homo, inliers
