"""Writes a synthetic dataset in the reference's on-disk layout (src/dataset_io.cpp) — TEST INFRASTRUCTURE for the loader in
global-lvba_b200/host/lvba_dataset.hpp and the offline tool tools/lvba_offline.cpp:

    <root>/all_pcd_body/<timestamp>.pcd  (pcl::PointXYZI; ascii / binary / binary_compressed in rotation)
    <root>/all_pcd_body/lidar_poses.txt  (TUM: t tx ty tz qx qy qz qw; with a comment, an empty and an unparsable line)

binary_compressed follows PCL: uint32 compressed size, uint32 raw size, LZF stream over the FIELD-major buffer.  The LZF
encoder below emits literals and back references (liblzf format), so the decoder's two paths are both exercised."""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np


def lzf_compress(data: bytes) -> bytes:
    """Greedy LZF: 3-byte hash table, back references up to 264 bytes at distance <= 8192, literal runs of <= 32."""
    n = len(data)
    out = bytearray()
    lit = bytearray()
    table = {}
    i = 0

    def flush():
        nonlocal lit
        while lit:
            chunk = lit[:32]
            out.append(len(chunk) - 1); out.extend(chunk)
            lit = lit[32:]

    while i < n:
        ref = table.get(data[i:i + 3]) if i + 2 < n else None
        if i + 2 < n:
            table[data[i:i + 3]] = i
        if ref is not None and 0 < i - ref <= 8192:
            length = 3
            while i + length < n and length < 264 and data[ref + length] == data[i + length]:
                length += 1
            flush()
            dist = i - ref - 1
            l2 = length - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8)); out.append(l2 - 7)
            out.append(dist & 0xff)
            i += length
        else:
            lit.append(data[i]); i += 1
    flush()
    return bytes(out)


def lzf_decompress(comp: bytes, out_len: int) -> bytes:
    out = bytearray()
    ip = 0
    while ip < len(comp):
        ctrl = comp[ip]; ip += 1
        if ctrl < 32:
            out.extend(comp[ip:ip + ctrl + 1]); ip += ctrl + 1
        else:
            length = ctrl >> 5
            if length == 7:
                length += comp[ip]; ip += 1
            dist = ((ctrl & 31) << 8 | comp[ip]) + 1; ip += 1
            for _ in range(length + 2):
                out.append(out[-dist])
    assert len(out) == out_len
    return bytes(out)


def write_pcd(path, xyz, intensity=None, encoding="binary"):
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    n = len(xyz)
    inten = np.zeros(n, np.float32) if intensity is None else np.asarray(intensity, np.float32)
    header = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
              f"COUNT 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {encoding}\n").encode()
    rec = np.column_stack([xyz, inten]).astype(np.float32)
    with open(path, "wb") as f:
        f.write(header)
        if encoding == "ascii":
            for r in rec:
                f.write((" ".join(repr(float(v)) for v in r) + "\n").encode())
        elif encoding == "binary":
            f.write(rec.tobytes())
        elif encoding == "binary_compressed":
            raw = rec.T.copy().tobytes()                                    # field-major
            comp = lzf_compress(raw)
            assert lzf_decompress(comp, len(raw)) == raw
            f.write(struct.pack("<II", len(comp), len(raw))); f.write(comp)
        else:
            raise ValueError(encoding)


def R_to_quat(R):
    """(w, x, y, z), w >= 0."""
    from scipy.spatial.transform import Rotation
    x, y, z, w = Rotation.from_matrix(R).as_quat()
    q = np.array([w, x, y, z])
    return -q if q[0] < 0 else q


def quat_to_R(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def write_lidar_dataset(root, scans, poses, t0=1000.0, dt=0.1, scale_quat=1.7):
    """scans: list of (n, 3) float32; poses (F, 12).  File names carry the timestamps; the TUM file carries DIFFERENT
    timestamps (the reference takes x_buf_[i].t from the file names, dataset_io.cpp:228-233) and un-normalised quaternions."""
    root = Path(root)
    d = root / "all_pcd_body"
    d.mkdir(parents=True, exist_ok=True)
    enc = ["binary", "ascii", "binary_compressed"]
    ts = [t0 + dt * i for i in range(len(scans))]
    for i, s in enumerate(scans):
        write_pcd(d / f"{ts[i]:.6f}.pcd", s, np.arange(len(s)) % 7, enc[i % 3])
    with open(d / "lidar_poses.txt", "w") as f:
        f.write("# timestamp tx ty tz qx qy qz qw\n\n")
        for i in range(len(scans)):
            q = R_to_quat(poses[i, :9].reshape(3, 3)) * scale_quat
            t = poses[i, 9:]
            f.write(f"{i * 1.0:.3f} {t[0]:.12f} {t[1]:.12f} {t[2]:.12f} {q[1]:.15f} {q[2]:.15f} {q[3]:.15f} {q[0]:.15f}\n")
            if i == 1:
                f.write("this line does not parse\n")
    return ts


# ------------------------------------------------------------------ the visual side of the dataset (tests of tools/lvba_offline --visual)
COLMAP_MAX_IMAGES = 2 ** 31 - 1


def image_name(ts):
    """LvbaSystem::getImagePath (src/lvba_system.cpp:2146-2148): std::to_string(double) prints six decimals."""
    return f"{ts:.6f}.png"


def write_image_set(root, image_ts, image_poses, extra_between=0):
    """<root>/all_image/<ts>.png (empty files: only the names are read) + image_poses.txt (TUM, T_W_I of the body at the image
    time).  extra_between: that many additional images / pose lines after every listed one, so that image_sample_step =
    extra_between + 1 selects exactly the listed ones (handleImages / handleCamPoses, src/dataset_io.cpp:77-131, :193-210)."""
    d = Path(root) / "all_image"
    d.mkdir(parents=True, exist_ok=True)
    with open(d / "image_poses.txt", "w") as f:
        f.write("# timestamp tx ty tz qx qy qz qw\n")
        for i, ts in enumerate(image_ts):
            for e in range(extra_between + 1):
                t = ts + 1e-3 * e
                (d / image_name(t)).write_bytes(b"")
                P = np.asarray(image_poses[i], np.float64)
                if e:                                                  # the skipped images carry poses that must not be used
                    P = P.copy(); P[9:] += 100.0
                q = R_to_quat(P[:9].reshape(3, 3))
                f.write(f"{t:.6f} {P[9]:.12f} {P[10]:.12f} {P[11]:.12f} {q[1]:.15f} {q[2]:.15f} {q[3]:.15f} {q[0]:.15f}\n")
    (d / "readme.txt").write_text("not an image")


def write_colmap_db(path, image_ts, keypoints, pair_matches, db_ids=None, kp_cols=4, full_path_names=False):
    """A COLMAP database with the three tables LvbaSystem::loadFromColmapDB reads (src/lvba_system.cpp:510-685).
    keypoints: list of (k_i, 2) float arrays; pair_matches: {(i, j): (m, 2) int array} with i < j dataset indices and columns
    (keypoint of i, keypoint of j); db_ids: the database image_id of every dataset image (default i + 1) — when db_ids[i] >
    db_ids[j] the stored columns are swapped, as COLMAP stores them by ascending image_id."""
    import sqlite3
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    if path.exists():
        path.unlink()
    n = len(image_ts)
    db_ids = list(range(1, n + 1)) if db_ids is None else [int(x) for x in db_ids]
    con = sqlite3.connect(str(path))
    con.execute("CREATE TABLE images (image_id INTEGER PRIMARY KEY, name TEXT NOT NULL UNIQUE, camera_id INTEGER NOT NULL)")
    con.execute("CREATE TABLE keypoints (image_id INTEGER PRIMARY KEY, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB)")
    con.execute("CREATE TABLE two_view_geometries (pair_id INTEGER PRIMARY KEY, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB, config INTEGER)")
    for i in range(n):
        name = ("images/" if full_path_names else "") + image_name(image_ts[i])
        con.execute("INSERT INTO images VALUES (?, ?, 1)", (db_ids[i], name))
        kp = np.asarray(keypoints[i], np.float32).reshape(-1, 2)
        blob = np.zeros((len(kp), kp_cols), np.float32)
        blob[:, :2] = kp
        if kp_cols > 2:
            blob[:, 2] = 1.5
        con.execute("INSERT INTO keypoints VALUES (?, ?, ?, ?)", (db_ids[i], len(kp), kp_cols, blob.tobytes()))
    for (i, j), m in pair_matches.items():
        m = np.asarray(m, np.uint32).reshape(-1, 2)
        a, b = db_ids[i], db_ids[j]
        if a > b:
            a, b = b, a
            m = m[:, ::-1]
        con.execute("INSERT INTO two_view_geometries VALUES (?, ?, 2, ?, 2)", (a * COLMAP_MAX_IMAGES + b, len(m), np.ascontiguousarray(m).tobytes()))
    con.commit()
    con.close()


def write_config_yaml(path, intr_full, width_full, height_full, scale, Rcl, Pcl, extrinsic_R=None, extrinsic_T=None, image_step=1,
                      db="Colmap/colmap.db", window=None, stage1_voxel=1.0, stage2_voxel=0.5, eigen1=(0.3, 0.1, 0.06, 0.03),
                      eigen2=(0.3, 0.1, 0.06, 0.03), anchor_leaf=0.05, lidar=True, visual=True):
    """The reference's config/config.yaml layout (multi-line flow list for Rcl included)."""
    R = np.asarray(Rcl, np.float64).reshape(3, 3)
    eR = np.eye(3) if extrinsic_R is None else np.asarray(extrinsic_R, np.float64).reshape(3, 3)
    eT = np.zeros(3) if extrinsic_T is None else np.asarray(extrinsic_T, np.float64)
    rows = ",\n      ".join(", ".join(f"{v:.12g}" for v in R[r]) for r in range(3))
    txt = f"""cam_model:
  cam_width: {width_full}
  cam_height: {height_full}
  scale: {scale}
  cam_fx: {intr_full[0]:.12g}
  cam_fy: {intr_full[1]:.12g}
  cam_cx: {intr_full[2]:.12g}
  cam_cy: {intr_full[3]:.12g}
  cam_d0: {intr_full[4]:.12g}
  cam_d1: {intr_full[5]:.12g}
  cam_d2: {intr_full[6]:.12g}
  cam_d3: {intr_full[7]:.12g}

extrin_calib:
  extrinsic_T: [{', '.join(f'{v:.12g}' for v in eT)}]
  extrinsic_R: [{', '.join(f'{v:.12g}' for v in eR.ravel())}]
  Rcl: [{rows}]
  Pcl: [{', '.join(f'{v:.12g}' for v in np.asarray(Pcl, np.float64))}]

data_config:
  data_path: "unused/"   # the tool takes --data
  colmap_db_path: "{db}"
  image_sample_step: {image_step}
  enable_lidar_ba: {'true' if lidar else 'false'}
  enable_visual_ba: {'true' if visual else 'false'}

window_ba:
  enable: {'true' if window else 'false'}
  size: {window or 10}
  anchor_leaf_size: {anchor_leaf}
  use_window_ba_rel: false

BALM_stage1:
  enable: true
  root_voxel_size: {stage1_voxel}
  eigen_ratio_array: [{', '.join(str(v) for v in eigen1)}]

BALM_stage2:
  root_voxel_size: {stage2_voxel}
  eigen_ratio_array: [{', '.join(str(v) for v in eigen2)}]

track_fusion:
  min_view_angle: 8.0 # degree
  reproj_mean_thr: 3.0

colmap_output:
  enable: true
  filter_size_points3D: 0.01
"""
    Path(path).write_text(txt)
