"""Drop-in for the reference's initialiser `estimate_camera_relative_poses` (template_handler.py:468-601) with the cost
evaluation on the GPU.

The reference scores C candidate target-pose tables -- one per camera: "the target poses as camera c alone sees them" --
by running `bundle_adjustment_costfn` over ALL observations once per table (compiled_helpers.py:517-549) and summing
the per-observation error norms per image (:550-560); the best candidate per image wins, and one more evaluation of the
winning table gives the per-image errors that `find_and_exclude_transform_outliers` consumes (:242-287).

Here the observation table is uploaded once (`GpuCostFn`), ALL C candidate tables are scored by one `pcs_costfn` call
with the per-image sums formed on the device, and the winning table by a second one: 2 calls instead of C + 1 passes.
Everything else -- the per-(camera, image) OpenCV pose estimates (`target_pose_in_cam_image`), the choice of the
reference pose, the transform algebra -- is the reference's own host logic, restated from the lines cited; the module
never imports pyCamSet and uses the target / detection / camera objects only through the methods the reference calls.

Return value: (Mrt_ac, Mat_rt, init_per_im_reproj_err) like the reference, INCLUDING its quirk that the third array has
2 M entries -- the per-image sums of the LAST candidate followed by those of the winning table (the reference appends
to the list it built for the last camera, :593-598).
"""
from __future__ import annotations

import numpy as np

from .handler import GpuCostFn


def _h_tform(points: np.ndarray, transform: np.ndarray) -> np.ndarray:
    """general_utils.h_tform (:236-260) for points (fill = 1)."""
    hp = np.concatenate([points, np.ones((len(points), 1))], axis=-1)[..., None]
    new = (transform[None, ...] @ hp)[..., 0]
    return new[:, :-1] / new[:, -1][..., None]


def _update_refpose(Mat_ac: np.ndarray, ref_pose: int) -> int:
    """check_feasiblity_and_update_refpose (template_handler.py:454-466): first pose every camera sees."""
    invisible = np.isnan(Mat_ac[:, :, 0, 0])
    visible_pose = ~np.any(invisible, axis=0)
    if not visible_pose[ref_pose]:
        f_index = int(np.argmax(visible_pose))
        if f_index == 0 and not visible_pose[0]:
            raise ValueError("Couldn't find an initial pose for all cameras.")
        ref_pose = f_index
    return ref_pose


def estimate_camera_relative_poses(calibration_target, detection, cams, ref_cam: int = 0, ref_pose: int = 0,
                                   max_bad_cams_iter: int = 10, device: int = 0):
    img_detections = detection.get_image_list()
    Mat_ac = np.array([[calibration_target.target_pose_in_cam_image(i, cam, mode="nan") for i in img_detections] for cam in cams])
    ref_pose = _update_refpose(Mat_ac, ref_pose)
    Mrt_ac = Mat_ac[:, ref_pose]
    Mac_rt = np.array([np.linalg.inv(m) for m in Mrt_ac])
    Mat_rt_ac = Mac_rt[:, None, ...] @ Mat_ac

    dists = np.array([cam.distortion_coefs for cam in cams]).squeeze()
    ints = np.array([cam.intrinsic for cam in cams])
    proj = ints @ Mrt_ac[:, :3, :]
    ps = calibration_target.point_data.reshape((-1, 3))
    target_shape = calibration_target.point_data.shape
    dd = detection.return_flattened_keys(target_shape[:-1]).get_data()
    n_cams, n_poses = Mat_ac.shape[0], int(detection.max_ims)

    tables = np.empty((n_cams, n_poses, ps.shape[0], 3))
    for c, Mat_rt_c in enumerate(Mat_rt_ac):
        nanform = np.isnan(Mat_rt_c[:, 0, 0])
        for idn, wasnan in enumerate(nanform):            # a missing pose estimate takes the previous image's (:523-529)
            if idn == 0 and wasnan:
                raise ValueError("No pose in first image")
            if wasnan:
                Mat_rt_c[idn] = Mat_rt_c[idn - 1]
        for m in range(n_poses):
            tables[c, m] = _h_tform(ps, Mat_rt_c[m])

    cost = GpuCostFn(dd, n_cams, n_poses, ps.shape[0], device=device)
    try:
        _, errors = cost.batch(tables, proj, ints, dists, errors=False)          # (C, M) per-image sums, all candidates at once
        estimate_locs = np.argmin(errors, axis=0)
        Mat_rt = np.array([Mt_rt_ac[e] for e, Mt_rt_ac in zip(estimate_locs, Mat_rt_ac.transpose((1, 0, 2, 3)))])
        best = np.array([_h_tform(ps, M) for M in Mat_rt])
        _, final = cost.batch(best[None], proj, ints, dists, errors=False)
    finally:
        cost.close()
    init_per_im_reproj_err = np.concatenate([errors[-1], final[0]])             # the reference's list re-use (:593-598)
    Mat_rt[ref_pose] = np.eye(4)
    return Mrt_ac, Mat_rt, init_per_im_reproj_err
