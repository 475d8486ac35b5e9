"""Pins the depth-rendering oracle (oracle/depth_oracle.py — buildGridMapFromOptimized + generateDepthWithVoxel,
src/lvba_system.cpp:1266-1338, 835-919) without a GPU: vectorised restatement == literal restatement, exactly."""
import numpy as np

from oracle import depth_oracle as dep
from oracle import synth


def _render(fn, s, **kw):
    return fn(s["scans"], s["poses"], s["frame_ts"], s["cams"], s["image_ts"], s["intr"], s["width"], s["height"], **kw)


def test_vectorised_equals_literal():
    for seed in (1, 2):
        s = synth.make_depth_scene(seed)
        a = _render(dep.render, s)
        b = _render(dep.render_literal, s)
        assert np.array_equal(a, b)
        filled = np.count_nonzero(a) / a.size
        assert 0.02 < filled < 0.9                                   # the cameras do see the scene, sparsely
        assert a[a > 0].min() >= 1e-3


def test_time_window_selects_frames():
    s = synth.make_depth_scene(3, F=8, M=3)
    full = _render(dep.render, s)
    narrow = _render(dep.render, s, half_window=0.05)
    assert np.count_nonzero(narrow) < np.count_nonzero(full)
    both = (narrow > 0) & (full > 0)
    assert np.all(full[both] <= narrow[both])                        # more points can only bring surfaces closer
    s2 = dict(s); s2["image_ts"] = s["image_ts"].copy(); s2["image_ts"][1] = np.nan
    img = _render(dep.render, s2)
    assert not img[1].any() and np.array_equal(img[0], full[0])      # an unparsable image name gives an empty image (:1309-1314)
    far = dict(s); far["image_ts"] = s["image_ts"] + 1e4
    assert not _render(dep.render, far).any()                        # no frame within +-0.5 s


def test_grid_key_and_pixel_quirks():
    assert dep.grid_keys(np.array([[0.2, -0.2, -1.0], [0.5, -0.5, 0.49999999]])).tolist() == [[0, -1, -3], [1, -2, 1]]
    # (int) truncation: a projection at u = -0.4 lands in pixel 0, one at -1.2 is dropped
    intr = np.array([100.0, 100.0, 0.0, 0.0, 0, 0, 0, 0])
    cam = np.concatenate([np.eye(3).ravel(), np.zeros(3)])
    depth = np.zeros((4, 4), np.float32)
    pw = np.array([[-0.004, 0.0, 1.0], [-0.012, 0.0, 1.0], [0.0, 0.0, 2.0], [0.0, 0.0, 0.0005]])
    dep._splat(depth, pw, cam, intr, 4, 4)
    assert depth[0, 0] == 1.0 and np.count_nonzero(depth) == 1
