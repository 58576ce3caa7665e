"""Duck-typed stand-ins for the reference's handler objects (pyCamSet is not installed on the GPU box).

They expose exactly the attributes pycamset_b200.handler reads from a real handler -- `op_fun.function_blocks`,
`op_fun.build_param_list`, `bundlePrimitive`, `target.point_data`, `detection.return_flattened_keys(...).get_data()`,
`get_bundle_adjustment_inputs`, `get_initial_params`, `problem_opts` -- and are filled from a golden case that the
real reference produced (tests/golden/make_golden.py).  Class names matter: the chain is keyed by block class names
(abstract_function_blocks.py:297) and the stock x -> array mapping is recognised by its qualified name.
"""
from __future__ import annotations

import numpy as np


class projection: pass            # noqa: E701  (names mirror function_block_implementations.py)
class extrinsic3D: pass           # noqa: E701
class template_points: pass       # noqa: E701
class rigidTform3d: pass          # noqa: E701
class free_point: pass            # noqa: E701


class _OpFun:
    def __init__(self, blocks):
        self.function_blocks = [b() for b in blocks]

    def build_param_list(self, *args):   # abstract_function_blocks.py:669-681
        return np.concatenate([np.asarray(a).flatten() for a in args])


def _fill_flat(data, dest, mask):        # compiled_helpers.py:155-177
    dest[np.asarray(mask, bool)] = data


class _Primitive:
    """TemplateBundlePrimitive / StandardBundlePrimitive (template_handler.py:32-78, standard_bundle_handler.py:46-107)."""

    def __init__(self, intr, extr, poses, intr_unfixed, extr_unfixed, poses_unfixed, points=None, bdpt_unfixed=None):
        self.intr, self.extr, self.poses = intr, extr, poses
        self.intr_unfixed, self.extr_unfixed, self.poses_unfixed = intr_unfixed, extr_unfixed, poses_unfixed
        self.bundle_pts = points
        if bdpt_unfixed is not None:
            self.bdpt_unfixed = bdpt_unfixed
        ni, ne, npo = int(intr_unfixed.sum()), int(extr_unfixed.sum()), int(poses_unfixed.sum())
        self.intr_end = 9 * ni
        self.extr_end = self.intr_end + 6 * ne
        self.pose_end = self.extr_end + 6 * npo
        self.n_free = (ni, ne, npo)

    def return_bundle_primitives(self, params):
        ni, ne, npo = self.n_free
        _fill_flat(params[self.extr_end:self.pose_end].reshape(npo, 6), self.poses, self.poses_unfixed)
        _fill_flat(params[self.intr_end:self.extr_end].reshape(ne, 6), self.extr, self.extr_unfixed)
        _fill_flat(params[:self.intr_end].reshape(ni, 9), self.intr, self.intr_unfixed)
        if self.bundle_pts is None:
            return self.intr, self.extr, self.poses
        _fill_flat(params[self.pose_end:], self.bundle_pts, self.bdpt_unfixed)
        return self.intr, self.extr, self.poses, self.bundle_pts.reshape(-1, 3)


class _Target:
    def __init__(self, template):
        self.point_data = template


class _Detection:
    def __init__(self, dd):
        self._dd = dd

    def return_flattened_keys(self, shape):
        return self

    def get_data(self):
        return self._dd


class TemplateBundleHandler:
    """Stand-in for template_handler.TemplateBundleHandler built from a golden case."""

    def __init__(self, g):
        C, M = int(g["n_cams"]), int(g["n_poses"])
        p = np.array(g["param0"], np.float64)
        unf = np.asarray(g["unfixed"], bool)
        intr, extr, poses = p[:9 * C].reshape(C, 9).copy(), p[9 * C:15 * C].reshape(C, 6).copy(), p[15 * C:15 * C + 6 * M].reshape(M, 6).copy()
        iu, eu, pu = unf[:9 * C].reshape(C, 9)[:, 0], unf[9 * C:15 * C].reshape(C, 6)[:, 0], unf[15 * C:15 * C + 6 * M].reshape(M, 6)[:, 0]
        selfcal = int(g["chain"]) == 1
        pts = p[15 * C + 6 * M:].copy() if selfcal else None
        self.bundlePrimitive = _Primitive(intr, extr, poses, iu, eu, pu, pts, unf[15 * C + 6 * M:] if selfcal else None)
        blocks = (projection, extrinsic3D, rigidTform3d, free_point) if selfcal else (projection, extrinsic3D, template_points)
        self.op_fun = _OpFun(blocks)
        self.target = _Target(np.array(g["template"], np.float64))
        self.detection = _Detection(np.array(g["dd"], np.float64))
        self.problem_opts = {"verbosity": 0, "max_nfev": 100}
        self.initial_params = np.array(g["x"], np.float64)

    def get_bundle_adjustment_inputs(self, x, make_points=False):
        return self.bundlePrimitive.return_bundle_primitives(x)

    def get_initial_params(self):
        return self.initial_params

    def get_detection_data(self, flatten=False):
        return self.detection.get_data()


class SelfBundleHandler(TemplateBundleHandler):
    def get_bundle_adjustment_inputs(self, x, make_points=False):
        return self.bundlePrimitive.return_bundle_primitives(x)


class FocalInKiloPixels(TemplateBundleHandler):
    """A user subclass in the spirit of examples/extend_param_handler.py: it re-parametrises x (focal lengths are
    optimised in units of 1000 px), so only the handler itself can map x to the parameter string."""

    def get_bundle_adjustment_inputs(self, x, make_points=False):
        x = np.array(x, np.float64)
        n = self.bundlePrimitive.intr_end
        v = x[:n].reshape(-1, 9)
        v[:, 0] *= 1000.0
        v[:, 2] *= 1000.0
        return self.bundlePrimitive.return_bundle_primitives(x)

    def get_initial_params(self):
        x = self.initial_params.copy()
        v = x[:self.bundlePrimitive.intr_end].reshape(-1, 9)
        v[:, 0] /= 1000.0
        v[:, 2] /= 1000.0
        return x


class UnknownChainHandler(TemplateBundleHandler):
    def __init__(self, g):
        super().__init__(g)

        class my_custom_block: pass  # noqa: E701
        self.op_fun = _OpFun((projection, extrinsic3D, my_custom_block))
