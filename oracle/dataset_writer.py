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
