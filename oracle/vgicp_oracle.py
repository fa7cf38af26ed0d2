"""vgicp_oracle.py — CPU (numpy / scipy) restatement of fast_gicp::FastVGICP as the reference uses it (estimator.cpp:263-303).

*** TEST INFRASTRUCTURE ONLY. ***  Only tests/ may import this module; the product (mvil_fusion_b200/) never does.

PARITY STATUS: "parity unpinned" against the compiled fast_gicp — the reference links a prebuilt libfast_gicp.a that is not in
the tree (vils_estimator/CMakeLists.txt:73-74) and its headers need PCL + Eigen, neither of which is in this image.  The headers
ARE in the tree, and every function below follows them line by line (paths relative to
vils_estimator/src/lidar_functions/fast_gicp/include/fast_gicp).  Third-party pieces restated from their published behaviour:
pcl::search::KdTree::nearestKSearch (exact k-NN, FLANN L2_Simple float distances -> scipy cKDTree on the float coordinates),
Eigen::JacobiSVD (-> numpy.linalg.svd), Eigen::LDLT (-> numpy.linalg.solve), pcl::Registration::getFitnessScore.
What pins it (tests/test_vgicp_oracle.py, CPU): central differences of the error sum reproduce b (the gradient convention of
linearize), H is symmetric PSD, and align() recovers a known rigid motion of a synthetic room scan.
"""
import numpy as np
from scipy.spatial import cKDTree

K_CORRESPONDENCES = 20          # gicp/impl/fast_gicp_impl.hpp:24


def calculate_covariances(xyz, k=K_CORRESPONDENCES):
    """gicp/impl/fast_gicp_impl.hpp:240-298, RegularizationMethod::PLANE.  xyz: n x 3 float32.  Returns (n x 3 x 3 covariances, n x k indices)."""
    pts = np.asarray(xyz, np.float32)
    tree = cKDTree(pts.astype(np.float64))
    _, idx = tree.query(pts.astype(np.float64), k=k)
    covs = np.zeros((len(pts), 3, 3))
    for i in range(len(pts)):
        nb = pts[idx[i]].astype(np.float64).T                      # :256-259 neighbours as double columns
        nb = nb - nb.mean(axis=1, keepdims=True)                   # :261
        cov = nb @ nb.T / k                                        # :262
        U, _, Vt = np.linalg.svd(cov)                              # :272 JacobiSVD, singular values descending
        covs[i] = U @ np.diag([1.0, 1.0, 1e-3]) @ Vt               # :280-281, :294-295
    return covs, idx


def voxel_coord(p, res):
    """gicp/fast_vgicp_voxel.hpp:150-152."""
    return tuple(np.floor(np.asarray(p, np.float64) / res - 0.5).astype(np.int64))


def create_voxelmap(xyz, covs, res):
    """gicp/fast_vgicp_voxel.hpp:113-148 with AdditiveGaussianVoxel (:92-107).  Returns {coord: [mean, cov, num_points, first_index]}."""
    vox = {}
    for i, p in enumerate(np.asarray(xyz, np.float32).astype(np.float64)):
        c = voxel_coord(p, res)
        v = vox.get(c)
        if v is None:
            v = vox[c] = [np.zeros(3), np.zeros((3, 3)), 0, i]
        v[0] = v[0] + p; v[1] = v[1] + covs[i]; v[2] += 1       # append
    for v in vox.values():
        v[0] = v[0] / v[2]; v[1] = v[1] / v[2]                   # finalize
    return vox


OFFSETS = {
    1: [(0, 0, 0)],
    7: [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)],
    27: [(i - 1, j - 1, k - 1) for i in range(3) for j in range(3) for k in range(3)],
}


def skew(x):
    return np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]], float)


def so3_exp(omega):
    """so3/so3.hpp:56-76 followed by Eigen's Quaternion::toRotationMatrix."""
    th2 = float(omega @ omega)
    if th2 < 1e-10:
        q4 = th2 * th2
        im = 0.5 - 1.0 / 48.0 * th2 + 1.0 / 3840.0 * q4
        re = 1.0 - 1.0 / 8.0 * th2 + 1.0 / 384.0 * q4
    else:
        th = np.sqrt(th2)
        im = np.sin(0.5 * th) / th
        re = np.cos(0.5 * th)
    w, x, y, z = re, im * omega[0], im * omega[1], im * omega[2]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class FastVGICP:
    """gicp/impl/fast_vgicp_impl.hpp + gicp/impl/lsq_registration_impl.hpp (LM stepper)."""

    def __init__(self, resolution=1.0, neighbor_search=1):
        self.res = resolution; self.offsets = OFFSETS[neighbor_search]
        self.max_iterations = 64; self.rotation_epsilon = 2e-3; self.transformation_epsilon = 5e-4     # lsq_registration_impl.hpp:11-13
        self.lm_max_iterations = 10; self.lm_init_lambda_factor = 1e-9; self.lm_lambda = -1.0           # :17-19
        self.final_hessian = np.eye(6); self.converged = False; self.nr_iterations = 0

    def set_input(self, source_xyz, target_xyz):
        self.src = np.asarray(source_xyz, np.float32)[:, :3]; self.tgt = np.asarray(target_xyz, np.float32)[:, :3]
        self.src_cov, _ = calculate_covariances(self.src); self.tgt_cov, _ = calculate_covariances(self.tgt)
        self.vox = create_voxelmap(self.tgt, self.tgt_cov, self.res)
        self.tgt_tree = cKDTree(self.tgt.astype(np.float64))

    def update_correspondences(self, T):
        """fast_vgicp_impl.hpp:76-118."""
        R = T[:3, :3]; t = T[:3, 3]
        self.corr = []; self.mahal = []
        for i, a in enumerate(self.src.astype(np.float64)):
            c = voxel_coord(R @ a + t, self.res)
            for o in self.offsets:
                v = self.vox.get((c[0] + o[0], c[1] + o[1], c[2] + o[2]))
                if v is not None:
                    self.corr.append((i, v))
                    self.mahal.append(np.linalg.inv(v[1] + R @ self.src_cov[i] @ R.T))      # the 4 x 4 of :108-113 is block-diagonal
        return len(self.corr)

    def linearize(self, T, with_h=True):
        """fast_vgicp_impl.hpp:120-176."""
        self.update_correspondences(T)
        return self._sum(T, with_h)

    def compute_error(self, T):
        """fast_vgicp_impl.hpp:178-203: correspondences and fused covariances of the last linearize."""
        return self._sum(T, False)[0]

    def _sum(self, T, with_h):
        R = T[:3, :3]; t = T[:3, 3]
        H = np.zeros((6, 6)); b = np.zeros(6); err = 0.0
        for (i, v), M in zip(self.corr, self.mahal):
            ta = R @ self.src[i].astype(np.float64) + t
            e = v[0] - ta
            w = np.sqrt(v[2])
            err += w * (e @ M @ e)
            if with_h:
                J = np.hstack([skew(ta), -np.eye(3)])
                H += w * (J.T @ M @ J); b += w * (J.T @ M @ e)
        return err, H, b

    def is_converged(self, delta):
        """lsq_registration_impl.hpp:76-86."""
        r = np.abs(delta[:3, :3] - np.eye(3)).max() / self.rotation_epsilon
        t = np.abs(delta[:3, 3]).max() / self.transformation_epsilon
        return max(r, t) < 1

    def step_lm(self, x0):
        """lsq_registration_impl.hpp:123-166.  Returns (ok, x0, delta)."""
        y0, H, b = self.linearize(x0)
        self.n_linearize += 1; self.last_error = y0
        if self.lm_lambda < 0.0:
            self.lm_lambda = self.lm_init_lambda_factor * np.abs(np.diag(H)).max()
        nu = 2.0
        delta = np.eye(4)
        for _ in range(self.lm_max_iterations):
            d = np.linalg.solve(H + self.lm_lambda * np.eye(6), -b)
            delta = np.eye(4); delta[:3, :3] = so3_exp(d[:3]); delta[:3, 3] = d[3:]
            xi = delta @ x0
            yi = self.compute_error(xi)
            rho = (y0 - yi) / (d @ (self.lm_lambda * d - b))
            if rho < 0:
                if self.is_converged(delta):
                    return True, x0, delta
                self.lm_lambda = nu * self.lm_lambda; nu = 2 * nu
                continue
            x0 = xi; self.last_error = yi
            self.lm_lambda = self.lm_lambda * max(1.0 / 3.0, 1 - (2 * rho - 1) ** 3)
            self.final_hessian = H
            return True, x0, delta
        return False, x0, delta

    def align(self, guess=None):
        """lsq_registration_impl.hpp:52-74 (computeTransformation)."""
        x0 = np.eye(4) if guess is None else np.array(guess, np.float64)
        self.lm_lambda = -1.0; self.converged = False; self.n_linearize = 0; self.last_error = 0.0
        for i in range(self.max_iterations):
            if self.converged:
                break
            self.nr_iterations = i
            ok, x0, delta = self.step_lm(x0)
            if not ok:
                break
            self.converged = self.is_converged(delta)
        self.final = x0
        return x0

    def fitness_score(self):
        """pcl::Registration::getFitnessScore(): the source moved by the float final transformation, mean squared 1-NN distance (float)."""
        Tf = self.final.astype(np.float32)
        p = self.src
        q = ((Tf[:3, 0] * p[:, :1] + Tf[:3, 1] * p[:, 1:2]) + Tf[:3, 2] * p[:, 2:3]) + Tf[:3, 3]
        _, idx = self.tgt_tree.query(q.astype(np.float64), k=1)
        d = q - self.tgt[idx]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        return float(d2.astype(np.float64).sum() / len(p))


def room_scan(rng, n, pose=None, noise=0.01):
    """A LiDAR-like scan of a 16 x 12 x 3 m room (the room of SURVEY §8d) seen from `pose` (4 x 4, sensor -> world): n x 4 float32."""
    per = n // 6
    pts = []
    for axis, val in ((0, -8.0), (0, 8.0), (1, -6.0), (1, 6.0), (2, 0.0), (2, 3.0)):
        p = np.stack([rng.uniform(-8, 8, per), rng.uniform(-6, 6, per), rng.uniform(0, 3, per)], 1)
        p[:, axis] = val
        pts.append(p)
    w = np.concatenate(pts) + rng.normal(0, noise, (per * 6, 3))
    if pose is not None:
        Ri = pose[:3, :3].T
        w = (w - pose[:3, 3]) @ Ri.T
    return np.c_[w, rng.uniform(0, 100, len(w))].astype(np.float32)
