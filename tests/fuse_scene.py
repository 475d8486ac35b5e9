"""Synthetic input of the track fusion for the tests: cameras along a path looking at a cloud of 3-D points, keypoints =
distorted projections + pixel noise (float32), pairwise matches between the images that see a point (plus wrong matches that merge
or pollute components), per-keypoint depth candidates (true point + noise, some missing, some wrong)."""
import numpy as np

INTR = np.array([646.78472, 646.65775, 313.456795, 261.399612, -0.07616, 0.123001, -0.00113, 0.000251])


def project(cam, X, intr=INTR):
    R = cam[:9].reshape(3, 3); t = cam[9:]
    pc = R @ X + t
    if pc[2] <= 0.2:
        return None
    x, y = pc[0] / pc[2], pc[1] / pc[2]
    r2 = x * x + y * y
    rad = 1 + intr[4] * r2 + intr[5] * r2 * r2
    xd = x * rad + 2 * intr[6] * x * y + intr[7] * (r2 + 2 * x * x)
    yd = y * rad + intr[6] * (r2 + 2 * y * y) + 2 * intr[7] * x * y
    u, v = intr[0] * xd + intr[2], intr[1] * yd + intr[3]
    if not (0 <= u < 640 and 0 <= v < 512):
        return None
    return u, v


def make(n_images=14, n_points=120, seed=0, wrong=0.06, no_depth=0.25, bad_depth=0.08, px_noise=0.4):
    rng = np.random.default_rng(seed)
    cams = np.zeros((n_images, 12))
    for i in range(n_images):                                   # cameras on an arc, looking roughly along +z, yawed
        ang = 0.10 * (i - n_images / 2)
        c, s = np.cos(ang), np.sin(ang)
        Rwc = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])      # camera-to-world
        Cw = np.array([0.55 * i, 0.05 * rng.normal(), 0.1 * rng.normal()])
        Rcw = Rwc.T
        cams[i, :9] = Rcw.ravel(); cams[i, 9:] = -Rcw @ Cw
    pts = np.column_stack([rng.uniform(-1, 0.55 * n_images + 1, n_points), rng.uniform(-2, 2, n_points), rng.uniform(4, 12, n_points)])
    kps = [[] for _ in range(n_images)]            # per image: (u, v, point id or -1)
    seen = [[] for _ in range(n_points)]           # per point: (image, keypoint index)
    for p in range(n_points):
        for i in range(n_images):
            if rng.random() < 0.25:
                continue
            uv = project(cams[i], pts[p])
            if uv is None:
                continue
            kps[i].append((uv[0] + px_noise * rng.normal(), uv[1] + px_noise * rng.normal(), p))
            seen[p].append((i, len(kps[i]) - 1))
    for i in range(n_images):                                   # clutter keypoints nobody matches
        for _ in range(5):
            kps[i].append((rng.uniform(0, 640), rng.uniform(0, 512), -1))
    kp_ptr = np.concatenate([[0], np.cumsum([len(k) for k in kps])]).astype(np.int64)
    kp_uv = np.array([[u, v] for k in kps for (u, v, _) in k], np.float32).reshape(-1, 2)
    pid = np.array([p for k in kps for (_, _, p) in k], np.int64)
    pair = {}
    for p in range(n_points):
        obs = seen[p]
        for a in range(len(obs)):
            for b in range(a + 1, len(obs)):
                if rng.random() < 0.7:
                    (ia, ka), (ib, kb) = obs[a], obs[b]
                    pair.setdefault((ia, ib), []).append((ka, kb))
    for _ in range(int(wrong * sum(len(v) for v in pair.values())) + 1):      # wrong matches: random keypoints of two images
        ia, ib = sorted(rng.choice(n_images, 2, replace=False))
        if len(kps[ia]) and len(kps[ib]):
            pair.setdefault((int(ia), int(ib)), []).append((int(rng.integers(len(kps[ia]))), int(rng.integers(len(kps[ib])))))
    matches = []
    for (ia, ib) in sorted(pair):                               # the reference's visiting order: pairs (i < j) ascending, stored order inside
        for (ka, kb) in pair[(ia, ib)]:
            matches.append((ia, ka, ib, kb))
    matches.append((0, 10 ** 6, 1, 0))                          # out-of-range keypoint: skipped (:944-947)
    matches = np.array(matches, np.int32)
    n_kp = int(kp_ptr[-1])
    kp_Xw = np.zeros((n_kp, 3)); kp_valid = np.zeros(n_kp, np.uint8)
    for g in range(n_kp):
        if pid[g] < 0 or rng.random() < no_depth:
            continue
        kp_valid[g] = 1
        kp_Xw[g] = pts[pid[g]] + 0.01 * rng.normal(size=3)
        if rng.random() < bad_depth:
            kp_Xw[g] += rng.normal(0, 1.0, 3)                   # depth from the wrong surface
    return dict(kp_ptr=kp_ptr, kp_uv=kp_uv, matches=matches, cams=cams, intr=INTR.copy(), kp_Xw=kp_Xw, kp_valid=kp_valid, pts=pts)
