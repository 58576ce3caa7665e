"""Gauge transform of the self-calibration result: drop-in for
`SelfBundleHandler.apply_gauge_transform` (standard_bundle_handler.py:339-410).

After a self-calibration the solved target points live in an arbitrary similarity gauge (7 coordinates were pinned,
:151-158).  The reference maps them back onto the target model: a scale s from the point pairs that sit one square apart
in the model, a rigid transform (Kabsch / SVD, compiled_helpers.py:728-762) from the scaled points onto the model, and
the same similarity applied to the target poses (conjugation) and camera extrinsics (right multiplication by the inverse)
so that every reprojection is unchanged.  Intrinsics never change (scale invariance).

The only super-linear piece -- the K x K distance tables the reference builds with scipy's cdist to find the pairs --
runs on the GPU as one pair-walking reduction (csrc/pcs_gauge.cu: nothing O(K^2) is stored); the rest is a 3 x 3 SVD and
C + M 4 x 4 products, restated here in numpy.  Like the reference this works for `target.valid_map is True`; index-pair
valid maps call an undefined helper in the reference and are restated as plain pair distances.
"""
from __future__ import annotations

import ctypes as ct
import logging

import numpy as np

from . import _lib as L


def _rodrigues(rvec):
    import cv2
    return cv2.Rodrigues(np.asarray(rvec, np.float64).reshape(3))[0]


def _tform(rvec, t):
    """general_utils.make_4x4h_tform (:360-385), 'opencv' convention."""
    T = np.eye(4)
    T[:3, :3] = _rodrigues(rvec)
    T[:3, 3] = np.asarray(t, np.float64).reshape(3)
    return T


def _to_rod(T):
    """general_utils.ext_4x4_to_rod (:262-272)."""
    import cv2
    return cv2.Rodrigues(np.ascontiguousarray(T[:3, :3]))[0].squeeze(), T[:3, 3]


def rigid_transform(v0, v1):
    """compiled_helpers.n_estimate_rigid_transform (:728-762): R, t with R v0 + t ~ v1."""
    t0, t1 = v0.mean(axis=0), v1.mean(axis=0)
    u, _, vh = np.linalg.svd((v0 - t0).T @ (v1 - t1))
    d = np.eye(3)
    d[-1, -1] = np.linalg.det(vh.T @ u.T)
    R = vh.T @ d @ u.T
    return R, -R @ t0 + t1


def gauge_scale(point_estimate, ref_points, visible, square_size, valid_map=True, device=0, rtol=1e-5, atol=1e-8):
    """s = mean(d_ref / d_estimate) over the valid pairs (:356-376)."""
    est = np.ascontiguousarray(point_estimate, np.float64).reshape(-1, 3)
    ref = np.ascontiguousarray(ref_points, np.float64).reshape(-1, 3)
    if isinstance(valid_map, bool):
        if not valid_map:
            raise ValueError("Target has given a valid map of False, which indicates no distance comparisons are valid.")
        vis = np.ascontiguousarray(visible, np.uint8)
        s, n = ct.c_double(0.0), ct.c_int64(0)
        L.check(L.load().pcs_gauge_scale(device, est.shape[0], est.ctypes.data, ref.ctypes.data, vis.ctypes.data, float(square_size),
                                         rtol, atol, ct.byref(s), ct.byref(n)))
        return s.value / n.value if n.value else float("nan")
    pairs = np.asarray(valid_map)[:, :2].astype(np.int64)
    new = np.linalg.norm(est[pairs[:, 0]] - est[pairs[:, 1]], axis=1)
    old = np.linalg.norm(ref[pairs[:, 0]] - ref[pairs[:, 1]], axis=1)
    return float(np.mean(old / new))


def apply_gauge_transform(proj, extr, poses, point_estimate, ref_points, visible, square_size, valid_map=True, device=0):
    """Returns (proj, extr, poses, new_points); extr / poses are updated in place like the reference does."""
    point_estimate = np.asarray(point_estimate, np.float64).reshape(-1, 3)
    ref_points = np.asarray(ref_points, np.float64).reshape(-1, 3)
    vm = np.asarray(visible, bool)
    s = gauge_scale(point_estimate, ref_points, vm, square_size, valid_map, device)
    new_points = s * point_estimate
    try:
        R, t = rigid_transform(new_points[vm], ref_points[vm])
        update = np.eye(4)
        update[:3, :3], update[:3, 3] = R, t
    except Exception as e:  # the reference's fallback (:383-387)
        logging.critical("Failed to find an acceptable gauge transform, returning the identity")
        logging.critical(f"Gave error: {e}")
        update = np.eye(4)
    inv_update = np.linalg.inv(update)
    new_points = (new_points @ update[:3, :3].T + update[:3, 3])
    for i in range(len(poses)):
        poses[i][3:] = poses[i][3:] * s
        poses[i][:3], poses[i][3:] = _to_rod(update @ _tform(poses[i][:3], poses[i][3:]) @ inv_update)
    for i in range(len(extr)):
        extr[i][3:] = extr[i][3:] * s
        extr[i][:3], extr[i][3:] = _to_rod(_tform(extr[i][:3], extr[i][3:]) @ inv_update)
    return proj, extr, poses, new_points


def apply_gauge_transform_for(handler, proj, extr, poses, point_estimate, device=0):
    """The same call with everything taken from a (self-calibration) reference handler, exactly as its own method reads it:
    target.point_data, target.valid_map, target.square_size, handler.visible_feature_mask."""
    t = handler.target
    return apply_gauge_transform(proj, extr, poses, point_estimate, np.asarray(t.point_data, np.float64).reshape(-1, 3),
                                 handler.visible_feature_mask, t.square_size, getattr(t, "valid_map", True), device)
