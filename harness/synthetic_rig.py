"""
Synthetic 2-camera rig and correspondence generator for the triangulation bench / parity tests.

Host-side (NumPy) mirror of the data path of the reference's accuracy harness
Work/triangulation_comparison/triangulation_comparison.py:
    Camera.camera_intrinsics   :94-107      Camera.camera_pose     :109-123
    Camera.project_points      :127-147     Camera.apply_noise     :149-162
    Camera.normalized_points   :164-173     finite_3D_points       :21-34
    cam_trajectory             :323-353     default_params         :266-287
    reset_random / rseed       :355-363, 370
plus a scalable cloud (uniform in the radius-r ball) for the 10 M / 100 M point workloads, which the
reference's 257-point integer lattice cannot provide.

Nothing here runs on the GPU; it only produces the (u1, P1, u2, P2) batches fed to the solvers.
"""
from math import asin, cos, sin
import numpy as np

RSEED = 123456789           # triangulation_comparison.py:370

DEFAULT_PARAMS = {          # triangulation_comparison.py:266-287
    "3D_points_r": 4,
    "cam_resolution": (640, 480),
    "cam_k1": 0.3,
    "cam_pose_offset": 40.,
    "cam_noise_sigma": 0.8,
    "cam_noise_discretized": True,
    "cam2_pose_sideways": 5.,
}


def P_from_R_and_t(R, t):
    """4x4 [R|t; 0 0 0 1]  (Work/python_libs/transforms.py:156-168)."""
    P = np.eye(4)
    P[0:3, 0:3] = R
    P[0:3, 3] = np.asarray(t, dtype=np.float64).reshape(3)
    return P


def rot_y(angle):
    """cv2.Rodrigues((0, angle, 0)) in closed form."""
    c, s = cos(angle), sin(angle)
    return np.array([[c, 0., s], [0., 1., 0.], [-s, 0., c]])


def finite_3D_points(r):
    """Integer lattice inside the radius-r sphere, homogeneous (triangulation_comparison.py:21-34)."""
    g = np.arange(-r, r + 1)
    x, y, z = np.meshgrid(g, g, g, indexing='ij')
    pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.float64)
    pts = pts[(pts ** 2).sum(axis=1) <= r * r]
    return np.concatenate([pts, np.ones((len(pts), 1))], axis=1)


def ball_3D_points(n, r, rng):
    """n points uniform in the radius-r ball (continuous analogue of finite_3D_points), homogeneous."""
    out = np.empty((n, 4))
    filled = 0
    while filled < n:
        m = int((n - filled) * 2.0) + 16
        c = rng.uniform(-r, r, size=(m, 3))
        c = c[(c ** 2).sum(axis=1) <= r * r][:n - filled]
        out[filled:filled + len(c), 0:3] = c
        filled += len(c)
    out[:, 3] = 1.
    return out


class Camera:
    def camera_intrinsics(self, resolution, k1=0.):
        f = min(resolution)
        c = np.array(resolution) / 2.
        K = np.eye(3)
        K[0, 0] = K[1, 1] = f
        K[0:2, 2] = c
        self.f, self.c, self.K = f, c, K
        self.dist_coeffs = np.array([k1, 0., 0., 0.])

    def camera_pose(self, offset, sideways=0., towards=0., angle=0.):
        cam_center = np.array([sideways, 0., -offset + towards])
        R = rot_y(angle)
        self.P = P_from_R_and_t(R, -R.dot(cam_center))[0:3, :]

    def project_points(self, points_3D, save_result=True):
        Xc = points_3D.dot(self.P.T)
        x = Xc[:, 0] / Xc[:, 2]; y = Xc[:, 1] / Xc[:, 2]
        k1 = self.dist_coeffs[0]
        if k1:
            rad = 1 + k1 * (x * x + y * y)
            x = x * rad; y = y * rad
        points_2D = np.stack([self.K[0, 0] * x + self.K[0, 2], self.K[1, 1] * y + self.K[1, 2]], axis=1)
        if save_result:
            self.points_2D_exact = self.points_2D = points_2D
        else:
            return points_2D

    def apply_noise(self, sigma, discretized=False, rng=np.random):
        if sigma:
            points_2D = self.points_2D_exact + rng.normal(0, sigma, self.points_2D_exact.shape)
        else:
            points_2D = self.points_2D_exact
        if discretized:
            points_2D = np.rint(points_2D)
        self.points_2D = points_2D

    def normalized_points(self):
        if not self.dist_coeffs[0]:
            u = np.array(self.points_2D)
            u[:, 0] -= self.c[0]
            u[:, 1] -= self.c[1]
            return u / self.f
        import cv2      # only the golden-file replay (k1 = 0.3) needs OpenCV's iterative undistortion
        return cv2.undistortPoints(np.array([self.points_2D]), self.K, self.dist_coeffs).reshape(-1, 2)


def cam_trajectory(cam_pose_offset, num_poses, from_sideways=0., to_sideways=0., from_towards=0., to_towards=0.,
                   from_angle=0., to_angle=0., angle_by_sideways=False):
    if angle_by_sideways:
        angle_values = np.linspace(asin(from_sideways / cam_pose_offset), asin(to_sideways / cam_pose_offset), num_poses)
        sideways_values = cam_pose_offset * np.sin(angle_values)
        towards_values = cam_pose_offset * (1 - np.cos(angle_values))
    else:
        sideways_values = np.linspace(from_sideways, to_sideways, num_poses)
        towards_values = np.linspace(from_towards, to_towards, num_poses)
        angle_values = np.linspace(from_angle, to_angle, num_poses)
    return {"sideways_values": sideways_values, "towards_values": towards_values, "angle_values": angle_values}


def default_trajectories(offset=40., num_poses=40, max_sideways=12., max_towards=12.):
    """The five 2nd-camera trajectories of triangulation_comparison.py:385-401."""
    return [
        cam_trajectory(offset, num_poses, to_sideways=max_sideways),
        cam_trajectory(offset, num_poses, to_towards=max_towards),
        cam_trajectory(offset, num_poses, from_sideways=max_sideways, to_sideways=max_sideways, to_towards=max_towards),
        cam_trajectory(offset, num_poses, to_sideways=max_sideways, angle_by_sideways=True),
        cam_trajectory(offset, num_poses, from_sideways=max_sideways, to_sideways=offset, angle_by_sideways=True),
    ]


# Named 2nd-camera poses (sideways, towards, angle) used by bench.py and the tests.
RIGS = {
    "translating": (5., 0., 0.),                                              # default_params: sideways 5
    "rotating": (12., 40. * (1 - cos(asin(12. / 40.))), asin(12. / 40.)),     # trajectory-4 end pose
    "forward": (0., 12., 0.),                                                 # trajectory-2 end pose (epipole in view)
    "general": (12., 12., 0.),                                                # trajectory-3 end pose
}


def make_cameras(rig="translating", k1=0., offset=40., resolution=(640, 480)):
    sideways, towards, angle = RIGS[rig] if isinstance(rig, str) else rig
    cam1, cam2 = Camera(), Camera()
    cam1.camera_pose(offset)
    cam2.camera_pose(offset, sideways, towards, angle)
    for cam in (cam1, cam2):
        cam.camera_intrinsics(resolution, k1)
    return cam1, cam2


def make_correspondences(n, rig="translating", sigma=0.8, discretized=False, seed=RSEED, r=4., dtype=np.float64,
                         return_cameras=False):
    """
    Seeded synthetic batch: cloud first, then cam-1 noise, then cam-2 noise (RandomState(seed)).
    Returns u1 (n,2), P1 (3,4), u2 (n,2), P2 (3,4), X (n,3) ground truth.  FP32 mode = same arrays rounded.
    """
    rng = np.random.RandomState(seed)
    cam1, cam2 = make_cameras(rig)
    X = ball_3D_points(n, r, rng)
    for cam in (cam1, cam2):
        cam.project_points(X)
        cam.apply_noise(sigma, discretized, rng)
    u1 = np.ascontiguousarray(cam1.normalized_points().astype(dtype))
    u2 = np.ascontiguousarray(cam2.normalized_points().astype(dtype))
    out = (u1, cam1.P.copy(), u2, cam2.P.copy(), np.ascontiguousarray(X[:, 0:3]))
    return out + (cam1, cam2) if return_cameras else out


BENCH_BASE_POINTS = 2_000_000


def bench_batch(n, rig="rotating", rank=0, sigma=0.8, base_points=BENCH_BASE_POINTS):
    """
    The batch bench.py times (and tests/test_gpu_parity.py::test_bench_input_parity checks against the oracle): one seeded
    batch of min(n, base_points) correspondences (seed RSEED + rank), tiled to n -- values repeat, which does not change
    the per-point cost, and the arrays (32 B/point) exceed the 126 MB L2 from 4 M points on.
    Returns u1 (n,2), P1, u2 (n,2), P2 and the number of distinct points.
    """
    base_n = min(n, base_points)
    u1b, P1, u2b, P2, _ = make_correspondences(base_n, rig, sigma=sigma, seed=RSEED + rank)
    reps = -(-n // base_n) if base_n else 1
    if reps == 1:
        return u1b, P1, u2b, P2, base_n
    return np.tile(u1b, (reps, 1))[:n], P1, np.tile(u2b, (reps, 1))[:n], P2, base_n


def circle_cameras(num_cams=8, offset=40., max_angle=asin(1.0) * 0.5):
    """num_cams poses on the trajectory-4/5 circle facing the cloud (multi-quadrotor scene, 3x4 each)."""
    Ps = []
    for a in np.linspace(0., max_angle, num_cams):
        cam = Camera()
        cam.camera_pose(offset, offset * sin(a), offset * (1 - cos(a)), a)
        Ps.append(cam.P.copy())
    return Ps


def make_multiview(n, num_cams=8, sigma=0.8, seed=RSEED, r=4., dtype=np.float64, p_visible=1.0):
    """
    Multi-quadrotor scene for the m-view solver: the cloud seen by `num_cams` cameras on the trajectory-4/5 circle.
    Returns us (m,n,2) normalised observations (pixel noise sigma / f), Ps (list of m 3x4), X (n,3), valid (m,n) bool
    (each view observes a point with probability p_visible; all True for p_visible >= 1).
    RNG order: cloud, then per camera its noise, then the visibility draws.
    """
    rng = np.random.RandomState(seed)
    X = ball_3D_points(n, r, rng)
    Ps = circle_cameras(num_cams)
    us = np.empty((num_cams, n, 2), dtype=dtype)
    for v, P in enumerate(Ps):
        cam = Camera()
        cam.P = P
        cam.camera_intrinsics((640, 480))
        cam.project_points(X)
        cam.apply_noise(sigma, False, rng)
        us[v] = cam.normalized_points().astype(dtype)
    valid = np.ones((num_cams, n), dtype=bool) if p_visible >= 1.0 else rng.uniform(size=(num_cams, n)) < p_visible
    return us, Ps, np.ascontiguousarray(X[:, 0:3]), valid
