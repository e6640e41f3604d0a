"""INRIA-3DGS .ply I/O in numpy (the on-disk format either side of the path).

Layout follows what the reference's loader expects (app/gaussians.cpp:75-171): binary little-endian
float32 properties x y z [nx ny nz] f_dc_0..2 f_rest_0..44 opacity scale_0..2 rot_0..3, with f_rest
channel-major on disk and coefficient-major / RGB-interleaved in memory, opacity as a logit, scale as
a log, rotation un-normalised (r,x,y,z).
"""
from __future__ import annotations

import numpy as np

_NAMES = (["x", "y", "z", "nx", "ny", "nz"] + ["f_dc_%d" % i for i in range(3)] + ["f_rest_%d" % i for i in range(45)]
          + ["opacity"] + ["scale_%d" % i for i in range(3)] + ["rot_%d" % i for i in range(4)])


def write_gs_ply(path: str, pos, sh, logit_opacity, log_scale, raw_rot) -> None:
    """sh is [P,16,3] (memory layout); the other arrays are the pre-activation values."""
    pos = np.asarray(pos, np.float32)
    P = pos.shape[0]
    sh = np.asarray(sh, np.float32).reshape(P, 16, 3)
    rows = np.zeros((P, len(_NAMES)), np.float32)
    rows[:, 0:3] = pos
    rows[:, 6:9] = sh[:, 0, :]
    # f_rest_i = feature[1 + i % 15][i // 15]  -> channel-major on disk
    rows[:, 9:54] = sh[:, 1:, :].transpose(0, 2, 1).reshape(P, 45)
    rows[:, 54] = np.asarray(logit_opacity, np.float32)
    rows[:, 55:58] = np.asarray(log_scale, np.float32)
    rows[:, 58:62] = np.asarray(raw_rot, np.float32)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P).encode())
        f.write("".join("property float %s\n" % n for n in _NAMES).encode())
        f.write(b"end_header\n")
        f.write(rows.astype("<f4").tobytes())


def read_gs_ply(path: str):
    """Returns post-activation (pos, scale, rotq, opacity, sh[P,16,3]) like GaussiansData."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        names, count, fmt = [], 0, None
        while True:
            line = f.readline().decode().strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element vertex"):
                count = int(line.split()[2])
            elif line.startswith("property"):
                t, n = line.split()[1:3]
                assert t in ("float", "float32"), "only float32 properties are supported"
                names.append(n)
            elif line == "end_header":
                break
        assert fmt == "binary_little_endian"
        data = np.frombuffer(f.read(count * len(names) * 4), "<f4").reshape(count, len(names))
    col = {n: i for i, n in enumerate(names)}
    pos = np.stack([data[:, col[k]] for k in "xyz"], axis=1).astype(np.float32)
    sh = np.zeros((count, 16, 3), np.float32)
    for c in range(3):
        sh[:, 0, c] = data[:, col["f_dc_%d" % c]]
    for i in range(45):
        sh[:, 1 + i % 15, i // 15] = data[:, col["f_rest_%d" % i]]
    opacity = (np.float32(1.0) / (np.float32(1.0) + np.exp(-data[:, col["opacity"]], dtype=np.float32))).astype(np.float32)
    scale = np.exp(np.stack([data[:, col["scale_%d" % i]] for i in range(3)], axis=1), dtype=np.float32)
    rot = np.stack([data[:, col["rot_%d" % i]] for i in range(4)], axis=1).astype(np.float32)
    rotq = (rot / np.sqrt((rot * rot).sum(axis=1, dtype=np.float32), dtype=np.float32)[:, None]).astype(np.float32)
    return pos, scale, rotq, opacity, sh
