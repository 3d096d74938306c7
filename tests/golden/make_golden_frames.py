"""Golden vectors of the camera-path generators, produced by the reference's own functions in the
build container: load_blender.pose_spherical (restated in make_golden.py because load_blender imports `magic`; checked
here against the source text's formula through load_llff's siblings), load_llff.poses_avg / render_path_spiral and the
`min_eval_frames` interpolation of load_llff._load_data (:73-78, scipy interp1d).

    python tests/golden/make_golden_frames.py        # writes tests/golden/stage_camera_paths.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (shims + reference on the path)
import load_llff  # noqa: E402  (the reference module)
from scipy.interpolate import interp1d  # noqa: E402


def main():
    rng = np.random.default_rng(3)
    args = np.array([[30.0, -30.0, 4.0], [-180.0, -30.0, 4.0], [123.4, -12.5, 2.75], [0.0, 0.0, 1.0]])
    spherical = np.stack([G.pose_spherical(*a) for a in args], 0)
    orbit = np.stack([G.pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, 40 + 1)[:-1]], 0).astype(np.float32)
    # LLFF-shaped poses [N,3,5]: rotations near identity, centres scattered, hwf column
    n = 9
    poses = np.zeros((n, 3, 5))
    for i in range(n):
        q, _ = np.linalg.qr(np.eye(3) + 0.1 * rng.standard_normal((3, 3)))
        poses[i, :, :3] = q * np.sign(np.diag(q))
        poses[i, :, 3] = 0.3 * rng.standard_normal(3)
        poses[i, :, 4] = [756.0, 1008.0, 800.0]
    c2w = load_llff.poses_avg(poses)
    up = load_llff.normalize(poses[:, :3, 1].sum(0))
    rads = np.percentile(np.abs(poses[:, :3, 3]), 90, 0)
    focal, zdelta = 3.7, 0.2
    spiral = np.stack(load_llff.render_path_spiral(c2w, up, rads, focal, zdelta, zrate=0.5, rots=2, N=30), 0)
    # the min_eval_frames interpolation, verbatim from load_llff.py:73-78
    pose_rows = rng.standard_normal((n, 17))
    min_eval_frames = 30
    m = int(np.ceil(min_eval_frames / (len(pose_rows) - 1)) * (len(pose_rows) - 1) + 1)
    repeat = (m - 1) // (len(pose_rows) - 1)
    interp = interp1d(np.arange(len(pose_rows)), pose_rows, axis=0)(np.linspace(start=0, stop=len(pose_rows) - 1, num=m))
    interp[::repeat, :] = pose_rows
    G.npz("stage_camera_paths.npz", spherical_args=args, spherical=spherical, orbit40=orbit, llff_poses=poses, poses_avg=c2w,
          up=up, rads=rads, focal=focal, zdelta=zdelta, spiral=spiral, pose_rows=pose_rows, min_eval_frames=min_eval_frames,
          repeat=repeat, pose_rows_interp=interp)


if __name__ == "__main__":
    main()
