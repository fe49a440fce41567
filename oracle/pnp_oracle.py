"""ORACLE -- test infrastructure, NOT product code.

The pose solver the matcher feeds (reference ``src/utils/eval_utils.py:18-42`` ``ransac_PnP``) restated.  The reference
delegates the arithmetic to a third-party dependency: ``cv2.solvePnPRansac`` of opencv_python (pinned 4.4.0.46 in the
reference's requirements.txt:5; this image ships 4.13.0).  OpenCV's algorithm (modules/calib3d/src/solvepnp.cpp): RANSAC over
5-point EPnP models with a fixed-seed RNG, inliers = reprojection error < ``reprojectionError`` pixels, adaptive iteration
count from ``confidence`` = 0.99 capped at ``iterationsCount``, final EPnP refit on the inlier set.  Because the sampling
sequence is OpenCV's own, parity of any other implementation is on the POSE (the reference's cm-degree metric,
src/evaluators/cmd_evaluator.py:11-33), not on bits.

Parity pin: ``tests/golden/pnp_scenes.npz`` holds what the UNMODIFIED reference function returns (imported from
/root/reference by tests/golden/make_golden.py) on seeded synthetic scenes; ``tests/test_oracle_golden.py`` checks this
restatement against it.  Used only by tests/ and bench.py's cpu_baseline leg.
"""
import numpy as np


def ransac_PnP(K, pts_2d, pts_3d, scale=1):
    """eval_utils.py:18-42, line by line."""
    import cv2
    dist_coeffs = np.zeros(shape=[8, 1], dtype="float64")                     # :20
    pts_2d = np.ascontiguousarray(pts_2d.astype(np.float64))                  # :22
    pts_3d = np.ascontiguousarray(pts_3d.astype(np.float64))                  # :23
    K = K.astype(np.float64)                                                  # :24
    pts_3d = pts_3d * scale                                                   # :26 (the reference scales in place)
    try:
        _, rvec, tvec, inliers = cv2.solvePnPRansac(pts_3d, pts_2d, K, dist_coeffs, reprojectionError=5,
                                                    iterationsCount=10000, flags=cv2.SOLVEPNP_EPNP)   # :28-29
        rotation = cv2.Rodrigues(rvec)[0]                                     # :31
        tvec = tvec / scale                                                   # :33
        pose = np.concatenate([rotation, tvec], axis=-1)                      # :34
        pose_homo = np.concatenate([pose, np.array([[0, 0, 0, 1]])], axis=0)  # :35
        inliers = [] if inliers is None else inliers                          # :37
        return pose, pose_homo, inliers
    except cv2.error:                                                         # :40-42
        return np.eye(4)[:3], np.eye(4), []


def pose_error(pose_pred, pose_gt):
    """(degrees, centimetres) -- cm-degree metric of cmd_evaluator.py:11-17 / eval_utils.py:45-63."""
    t_err = np.linalg.norm(pose_pred[:3, 3] - pose_gt[:3, 3]) * 100
    trace = min(np.trace(pose_pred[:3, :3] @ pose_gt[:3, :3].T), 3.0)
    return float(np.rad2deg(np.arccos(np.clip((trace - 1.0) / 2.0, -1.0, 1.0)))), float(t_err)
