"""Synthetic camera rigs + calibration-target observations.

The rig recipe follows the reference's example (examples/make_camera_ring.py:7-16): camera b of C has the
world->camera transform rvec = (0, 2 pi b / C, 0), t = (0, 0, 0.2): every camera looks at the origin from
0.2 m.  The target is the ChArUco(10, 10, 4) board of target_charuco.py:33-42 (81 inner corners, 4 mm pitch,
metres).  Everything here is host-side data preparation; tensors are produced with torch so that the 10^8
observation configuration can be generated directly in HBM, pose-chunk by pose-chunk.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


def charuco_points(nx: int = 10, ny: int = 10, square_mm: float = 4.0) -> np.ndarray:
    """Inner chessboard corners of a CharucoBoard((nx, ny), square) in metres, row-major (y outer, x inner);
    same values and order as cv2's board.getChessboardCorners() used by target_charuco.py:42."""
    s = square_mm / 1000.0
    pts = [((x + 1) * s, (y + 1) * s, 0.0) for y in range(ny - 1) for x in range(nx - 1)]
    return np.asarray(pts, dtype=np.float64)


def _rodrigues_t(r: torch.Tensor) -> torch.Tensor:
    """rvec (..., 3) -> R (..., 3, 3); theta < 1e-10 -> identity (compiled_helpers.py:197-235)."""
    theta = torch.linalg.norm(r, dim=-1, keepdim=True)
    small = theta < 1e-10
    th = torch.where(small, torch.ones_like(theta), theta)
    k = r / th
    kx, ky, kz = k.unbind(-1)
    zero = torch.zeros_like(kx)
    Kx = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], -1).reshape(*r.shape[:-1], 3, 3)
    ct = torch.cos(th)[..., None]
    st = torch.sin(th)[..., None]
    eye = torch.eye(3, dtype=r.dtype, device=r.device).expand(*r.shape[:-1], 3, 3)
    R = ct * eye + (1 - ct) * (k[..., :, None] * k[..., None, :]) + st * Kx
    return torch.where(small[..., None], eye, R)


def _log_so3(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> rvec (numpy, single matrix)."""
    c = (np.trace(R) - 1.0) / 2.0
    c = min(1.0, max(-1.0, c))
    theta = math.acos(c)
    if theta < 1e-12:
        return np.zeros(3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if math.pi - theta < 1e-6:  # near pi: use the symmetric part
        A = (R + np.eye(3)) / 2.0
        ax = np.sqrt(np.maximum(np.diag(A), 0.0))
        i = int(np.argmax(ax))
        ax = A[i] / ax[i]
        ax /= np.linalg.norm(ax)
        if np.dot(ax, w) < 0:
            ax = -ax
        return ax * theta
    return w / (2.0 * math.sin(theta)) * theta


def ring_extrinsics(n_cams: int, radius: float = 0.2) -> np.ndarray:
    """(C, 6) [rvec | t] of the reference's camera ring (examples/make_camera_ring.py:9-11)."""
    e = np.zeros((n_cams, 6))
    e[:, 1] = 2.0 * np.pi * np.arange(n_cams) / n_cams
    e[:, 5] = radius
    return e


def dome_extrinsics(n_cams: int, rng: np.random.Generator, r_min: float = 0.2, r_max: float = 0.4,
                    max_polar_deg: float = 75.0) -> np.ndarray:
    """(C, 6) cameras on a Fibonacci spherical cap (polar angle <= max_polar_deg about -z, the side the
    target's front face looks at in its reference pose), all looking at the origin."""
    e = np.zeros((n_cams, 6))
    golden = math.pi * (3.0 - math.sqrt(5.0))
    cos_max = math.cos(math.radians(max_polar_deg))
    for i in range(n_cams):
        cz = 1.0 - (1.0 - cos_max) * (i + 0.5) / n_cams  # cos(polar) uniformly spaced -> equal area
        sz = math.sqrt(max(0.0, 1.0 - cz * cz))
        phi = golden * i
        d = np.array([sz * math.cos(phi), sz * math.sin(phi), -cz])  # direction origin -> camera centre
        rad = rng.uniform(r_min, r_max)
        zc = -d  # camera z axis looks at the origin
        up = np.array([0.0, 1.0, 0.0]) if abs(zc[1]) < 0.95 else np.array([1.0, 0.0, 0.0])
        xc = np.cross(up, zc)
        xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        R = np.stack([xc, yc, zc])  # rows = camera axes in world coords -> world->camera rotation
        e[i, :3] = _log_so3(R)
        e[i, 3:] = -R @ (d * rad)  # = (0, 0, rad)
    return e


def default_intrinsics(n_cams: int) -> np.ndarray:
    """Reference default Camera (cameras/camera.py:20-24): f = 1000 px, pp = (500, 500), no distortion."""
    q = np.zeros((n_cams, 9))
    q[:, 0] = q[:, 2] = 1000.0
    q[:, 1] = q[:, 3] = 500.0
    return q


def perturbed_intrinsics(n_cams: int, rng: np.random.Generator) -> np.ndarray:
    """Radial + tangential distortion rig of SURVEY.md 8d (configs 4 / 5)."""
    q = np.zeros((n_cams, 9))
    f = rng.uniform(900.0, 1300.0, n_cams)
    q[:, 0] = f
    q[:, 2] = f * (1.0 + rng.normal(0.0, 1e-3, n_cams))
    q[:, 1] = 500.0 + rng.normal(0.0, 10.0, n_cams)
    q[:, 3] = 500.0 + rng.normal(0.0, 10.0, n_cams)
    q[:, 4] = rng.normal(0.0, 0.05, n_cams)   # k1
    q[:, 5] = rng.normal(0.0, 0.02, n_cams)   # k2
    q[:, 6] = rng.normal(0.0, 1e-3, n_cams)   # p1
    q[:, 7] = rng.normal(0.0, 1e-3, n_cams)   # p2
    q[:, 8] = 0.0                             # k3
    return q


def random_poses(n_poses: int, template: np.ndarray, rng: np.random.Generator, rot_sigma: float = 0.3,
                 t_sigma: float = 0.010) -> np.ndarray:
    """(M, 6) target poses.  Pose 0 is the identity (the handler fixes it, template_handler.py:134-137);
    the others rotate by rvec ~ N(0, rot_sigma^2) about the board centre and shift it by N(0, t_sigma^2)."""
    p = np.zeros((n_poses, 6))
    if n_poses > 1:
        p[1:, :3] = rng.normal(0.0, rot_sigma, (n_poses - 1, 3))
        centre = torch.as_tensor(template.mean(axis=0))
        R = _rodrigues_t(torch.as_tensor(p[1:, :3])).numpy()
        shift = rng.normal(0.0, t_sigma, (n_poses - 1, 3))
        p[1:, 3:] = centre.numpy() + shift - R @ centre.numpy()
    return p


def project_torch(intr: torch.Tensor, Xc: torch.Tensor) -> torch.Tensor:
    """Pinhole + Brown-Conrady forward model (function_block_implementations.py:27-47); broadcasts."""
    xn = Xc[..., 0] / Xc[..., 2]
    yn = Xc[..., 1] / Xc[..., 2]
    r2 = xn * xn + yn * yn
    fx, px, fy, py, k1, k2, p1, p2, k3 = intr.unbind(-1)
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xD = xn * rad + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
    yD = yn * rad + p1 * (r2 + 2 * yn * yn) + 2 * p2 * xn * yn
    return torch.stack([fx * xD + px, fy * yD + py], -1)


@dataclass
class SyntheticRig:
    """Structure-of-arrays observation table plus the ground-truth parameter arrays."""
    cam: torch.Tensor       # (N,) int32
    pose: torch.Tensor      # (N,) int32
    key: torch.Tensor       # (N,) int32
    uv: torch.Tensor        # (N, 2) float64
    template: np.ndarray    # (K, 3)
    intr: np.ndarray        # (C, 9) truth
    extr: np.ndarray        # (C, 6) truth
    poses: np.ndarray       # (M, 6) truth
    image_size: tuple = (1000.0, 1000.0)

    @property
    def n_obs(self) -> int:
        return int(self.cam.shape[0])

    @property
    def n_cams(self) -> int:
        return int(self.intr.shape[0])

    @property
    def n_poses(self) -> int:
        return int(self.poses.shape[0])

    @property
    def n_keys(self) -> int:
        return int(self.template.shape[0])

    def dd(self) -> np.ndarray:
        """N x 5 float64 [cam, img, key, u, v] table in the reference's layout (target_detections.py:51-55)."""
        out = np.empty((self.n_obs, 5))
        out[:, 0] = self.cam.cpu().numpy()
        out[:, 1] = self.pose.cpu().numpy()
        out[:, 2] = self.key.cpu().numpy()
        out[:, 3:] = self.uv.cpu().numpy()
        return out

    def param_string(self, intr=None, extr=None, poses=None, points=None) -> np.ndarray:
        parts = [self.intr if intr is None else intr, self.extr if extr is None else extr,
                 self.poses if poses is None else poses]
        if points is not None:
            parts.append(points)
        return np.concatenate([np.asarray(a, np.float64).ravel() for a in parts])

    def perturbed(self, rng: np.random.Generator, rel: float = 1e-3):
        """truth + rel * relative perturbation; pose 0 stays the identity."""
        intr = self.intr * (1 + rel * rng.normal(size=self.intr.shape)) + rel * 1e-2 * rng.normal(size=self.intr.shape) * (self.intr == 0)
        extr = self.extr * (1 + rel * rng.normal(size=self.extr.shape)) + rel * 1e-2 * rng.normal(size=self.extr.shape)
        poses = self.poses * (1 + rel * rng.normal(size=self.poses.shape)) + rel * 1e-2 * rng.normal(size=self.poses.shape)
        poses[0] = 0.0
        return intr, extr, poses


def make_rig(n_cams: int, n_poses: int, *, layout: str = "ring", distortion: bool = False, seed: int = 0,
             noise_px: float = 0.1, detect_prob: float = 1.0, pose_start: int = 0, pose_stop: int | None = None,
             device: str | torch.device = "cpu", pose_chunk: int = 256, order: str = "cam") -> SyntheticRig:
    """Generate observations of the ChArUco(10,10,4) board by a camera ring / dome.

    Visibility: point in front of the camera (z > 0.02), inside the 1000 x 1000 image and the
    board's front face turned towards the camera.  ``detect_prob`` < 1 additionally drops detections at
    random, which makes the per-(camera, pose) runs ragged the way real detections are.
    ``pose_start:pose_stop`` restricts generation to a pose range of the SAME global rig (a rank's shard):
    all random draws for the rig itself are made for the full pose count so shards are consistent.
    ``order`` = "cam" returns rows camera-major (the reference's detection order, camera_calibrator.py:293-317),
    "pose" returns them pose-major.
    """
    rng = np.random.default_rng(seed)
    template = charuco_points()
    K = template.shape[0]
    if layout == "ring":
        extr = ring_extrinsics(n_cams)
    elif layout == "dome":
        extr = dome_extrinsics(n_cams, rng)
    else:
        raise ValueError(f"unknown layout {layout!r}")
    intr = perturbed_intrinsics(n_cams, rng) if distortion else default_intrinsics(n_cams)
    poses = random_poses(n_poses, template, rng)
    pose_stop = n_poses if pose_stop is None else pose_stop

    dev = torch.device(device)
    f64 = dict(dtype=torch.float64, device=dev)
    T = torch.as_tensor(template, **f64)
    intr_t = torch.as_tensor(intr, **f64)
    extr_t = torch.as_tensor(extr, **f64)
    Rc = _rodrigues_t(extr_t[:, :3])                         # (C,3,3)
    tc = extr_t[:, 3:]
    normal = torch.tensor([0.0, 0.0, -1.0], **f64)           # board front face (seen by camera 0 at pose 0)
    gen = torch.Generator(device=dev)
    cams, pss, keys, uvs = [], [], [], []
    # Chunks are aligned to multiples of `pose_chunk` of the GLOBAL pose axis and always generated in full, then cut
    # to [pose_start, pose_stop): the random draws (detection mask, pixel noise) of a pose therefore do not depend on
    # how the poses are sharded, and the union of the shards of any world size is the same table.
    for c0 in range((pose_start // pose_chunk) * pose_chunk, pose_stop, pose_chunk):
        c1 = min(n_poses, c0 + pose_chunk)
        P = torch.as_tensor(poses[c0:c1], **f64)
        Rm = _rodrigues_t(P[:, :3])                          # (m,3,3)
        Xw = torch.einsum("mab,kb->mka", Rm, T) + P[:, None, 3:]          # (m,K,3)
        Xc = torch.einsum("cab,mkb->cmka", Rc, Xw) + tc[:, None, None, :]  # (C,m,K,3)
        uv = project_torch(intr_t[:, None, None, :], Xc)                   # (C,m,K,2)
        n_c = torch.einsum("cab,mb->cma", Rc, torch.einsum("mab,b->ma", Rm, normal))  # (C,m,3)
        facing = (n_c[:, :, None, :] * Xc).sum(-1) < 0
        vis = facing & (Xc[..., 2] > 0.02) & (uv[..., 0] >= 0) & (uv[..., 0] <= 1000.0) & (uv[..., 1] >= 0) & (uv[..., 1] <= 1000.0)
        gen.manual_seed(seed * 1000003 + c0)
        if detect_prob < 1.0:
            vis &= torch.rand(vis.shape, generator=gen, device=dev) < detect_prob
        noise = torch.randn(uv.shape, generator=gen, **f64) * noise_px
        uv = uv + noise
        m0, m1 = max(c0, pose_start), min(c1, pose_stop)
        if m0 >= m1:
            continue
        vis = vis[:, m0 - c0:m1 - c0]
        uv = uv[:, m0 - c0:m1 - c0]
        if order == "pose":
            vis = vis.permute(1, 0, 2)
            uv = uv.permute(1, 0, 2, 3)
            mi, ci, ki = torch.nonzero(vis, as_tuple=True)
            uvs.append(uv[mi, ci, ki])
        else:
            ci, mi, ki = torch.nonzero(vis, as_tuple=True)
            uvs.append(uv[ci, mi, ki])
        cams.append(ci.to(torch.int32)); pss.append((mi + m0).to(torch.int32)); keys.append(ki.to(torch.int32))
    cam = torch.cat(cams); pose = torch.cat(pss); key = torch.cat(keys); uvt = torch.cat(uvs)
    if order == "cam" and len(cams) > 1:
        # chunks are camera-major internally; make the whole table camera-major (stable)
        idx = torch.argsort(cam.to(torch.int64) * (n_poses + 1) + pose.to(torch.int64), stable=True)
        cam, pose, key, uvt = cam[idx], pose[idx], key[idx], uvt[idx]
    return SyntheticRig(cam=cam, pose=pose, key=key, uv=uvt.contiguous(), template=template, intr=intr, extr=extr,
                        poses=poses)
