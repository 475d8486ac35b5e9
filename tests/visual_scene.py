"""A small LiDAR + camera dataset on disk in the reference's layout, for the tests of `lvba_offline --visual`: the room corner of
oracle.synth.make_scan_scene seen by a camera riding on the LiDAR body (looking sideways at the wall y = -2.4 while moving along it), landmarks = world points of the
scans themselves (so they lie on the planes the voxel map finds), keypoints = distorted projections + pixel noise, inlier matches
between every pair of images that share a landmark, written into a COLMAP database."""
import numpy as np

from oracle import dataset_writer as dw, synth

# A small image on purpose: the depth candidate of a keypoint needs LiDAR depth in all four neighbouring pixels (fetchDepthBilinear,
# include/utils.hpp:246-275), so the scans must cover the image densely — 80 x 64 pixels after scaling keeps the point count modest.
WIDTH_FULL, HEIGHT_FULL = 160, 128
INTR_FULL = np.array([96.0, 96.3, 79.7, 64.2, -0.05, 0.01, 5e-4, -3e-4])
SCALE = 0.5
RCL = np.array([[-1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, -1.0, 0.0]])      # camera z = -body y (sideways, towards the wall y = -2.4), x = -body x, y = -body z
PCL = np.array([0.02, 0.10, -0.03])


def project(Rcw, tcw, X, intr, width, height):
    pc = Rcw @ X + tcw
    if pc[2] <= 0.3:
        return None
    x, y = pc[0] / pc[2], pc[1] / pc[2]
    r2 = x * x + y * y
    rad = 1 + intr[4] * r2 + intr[5] * r2 * r2
    xd = x * rad + 2 * intr[6] * x * y + intr[7] * (r2 + 2 * x * x)
    yd = y * rad + intr[6] * (r2 + 2 * y * y) + 2 * intr[7] * x * y
    u, v = intr[0] * xd + intr[2], intr[1] * yd + intr[3]
    if not (2 <= u < width - 2 and 2 <= v < height - 2):
        return None
    return u, v


def make(root, seed=3, W=8, n_per_scan=30000, n_landmarks=260, cam_noise=(0.01, 0.03), px_noise=0.1, extra_between=1, shuffle_db_ids=True):
    """Writes the dataset under `root` and returns a dict with the ground truth."""
    rng = np.random.default_rng(seed)
    scans, poses = synth.make_scan_scene(seed, W=W, n_per_scan=n_per_scan)
    ts = dw.write_lidar_dataset(root, scans, poses)                                   # frame timestamps 1000.0 + 0.1 i
    image_ts = [t + 0.013 for t in ts]                                               # one image just after every scan
    intr = INTR_FULL.copy(); intr[:4] *= SCALE
    width, height = int(round(WIDTH_FULL * SCALE)), int(round(HEIGHT_FULL * SCALE))
    # quaternion round trip of the pose file (15 digits): what the loader reads
    body = poses.copy()
    for i in range(W):
        body[i, :9] = dw.quat_to_R(dw.R_to_quat(poses[i, :9].reshape(3, 3))).ravel()
    cams_true = np.zeros((W, 12))
    for i in range(W):
        Rcw = RCL @ body[i, :9].reshape(3, 3).T
        cams_true[i, :9] = Rcw.ravel(); cams_true[i, 9:] = -Rcw @ body[i, 9:] + PCL
    # odometry image poses: the true body pose with a small error (image 0 exact: the solver holds camera 0 fixed)
    image_poses = body.copy()
    for i in range(1, W):
        image_poses[i, :9] = (body[i, :9].reshape(3, 3) @ synth.so3_exp(rng.normal(0, cam_noise[0], (1, 3)))[0]).ravel()
        image_poses[i, 9:] += rng.normal(0, cam_noise[1], 3)
    world = np.concatenate([np.asarray(s, np.float64) @ body[i, :9].reshape(3, 3).T + body[i, 9:] for i, s in enumerate(scans)])
    on_wall = np.abs(world[:, 1] + 2.4) < 0.03                                        # the wall the cameras look at (they move along it)
    cand = world[on_wall]
    pts = cand[rng.choice(len(cand), n_landmarks, replace=False)]
    kps = [[] for _ in range(W)]
    seen = [[] for _ in range(n_landmarks)]
    for p in range(n_landmarks):
        for i in range(W):
            uv = project(cams_true[i, :9].reshape(3, 3), cams_true[i, 9:], pts[p], intr, width, height)
            if uv is None or rng.random() < 0.15:
                continue
            kps[i].append((uv[0] + px_noise * rng.normal(), uv[1] + px_noise * rng.normal()))
            seen[p].append((i, len(kps[i]) - 1))
    for i in range(W):
        for _ in range(6):
            kps[i].append((rng.uniform(0, width), rng.uniform(0, height)))            # clutter nobody matches
    pair = {}
    for p in range(n_landmarks):
        o = seen[p]
        for a in range(len(o)):
            for b in range(a + 1, len(o)):
                if rng.random() < 0.8:
                    pair.setdefault((o[a][0], o[b][0]), []).append((o[a][1], o[b][1]))
    pair[(0, 1)] = pair.get((0, 1), []) + [(10 ** 6, 0)]                              # out-of-range keypoint: dropped by the reader (:668-672)
    keypoints = [np.array(k, np.float32).reshape(-1, 2) for k in kps]
    db_ids = list(range(1, W + 1))
    if shuffle_db_ids:
        db_ids = [int(x) for x in rng.permutation(np.arange(3, W + 3))]
    dw.write_image_set(root, image_ts, image_poses, extra_between=extra_between)
    dw.write_colmap_db(root / "Colmap" / "colmap.db", image_ts, keypoints, {k: np.array(v) for k, v in pair.items()}, db_ids=db_ids)
    dw.write_config_yaml(root / "config.yaml", INTR_FULL, WIDTH_FULL, HEIGHT_FULL, SCALE, RCL, PCL, image_step=extra_between + 1, lidar=False, stage2_voxel=0.5)
    return dict(scans=scans, poses=body, ts=ts, image_ts=image_ts, image_poses=image_poses, cams_true=cams_true, intr=intr, width=width, height=height,
                keypoints=keypoints, pair=pair, pts=pts, db_ids=db_ids)
