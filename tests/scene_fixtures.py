"""Scenes for the run-time scene API tests (B200PT_PROFILE_OPT_V4)."""
import numpy as np


def default_v4_scene():
    """The reference's InitializeScene data (demofox_path_tracing_optimization_v4.cpp:1403-1496) as arrays."""
    T = np.array([0.0, 0.0, 10.0], dtype=np.float32)
    f = np.float32
    quads = np.array([
        [[-25, -12.5, 5], [25, -12.5, 5], [25, -12.5, -5], [-25, -12.5, -5]],
        [[-25, -1.5, 5], [25, -1.5, 5], [25, -10.5, 5], [-25, -10.5, 5]],
        [[-7.5, 12.5, 5], [7.5, 12.5, 5], [7.5, 12.5, -5], [-7.5, 12.5, -5]],
        [[-5, 12.4, 2.5], [5, 12.4, 2.5], [5, 12.4, -2.5], [-5, 12.4, -2.5]]], dtype=np.float32)
    for i in (0, 2, 3):
        quads[i] = quads[i] + T  # the backdrop is not translated
    spheres = np.array([[f(-18.0) + f(6.0) * f(i), -8.0, 0.0 + 10.0, 2.8] for i in range(7)], dtype=np.float32)
    mats = np.zeros((11, 17), dtype=np.float32)
    mats[0, 0:3] = 0.7
    mats[1, 0:3] = 0.35
    mats[2, 0:3] = 0.7
    mats[3, 3:6] = np.array([1.0, 0.9, 0.7], dtype=np.float32) * f(20.0)
    for i in range(7):
        r = (f(i) / f(6)) * f(0.5)
        mats[4 + i] = [0.9, 0.25, 0.25, 0, 0, 0, 0.02, r, f(1.0) * f(0.8), f(1.0) * f(0.8), f(1.0) * f(0.8), 1.1, 1.0, r, 0.0, 0.5, 1.0]
    return quads, spheres, mats


def random_v4_scene(seed, nq=5, ns=6):
    """A scene 'of the named shape' with a different primitive count: floor + random quads, random spheres."""
    rng = np.random.default_rng(seed)
    quads = np.zeros((nq, 4, 3), dtype=np.float32)
    quads[0] = [[-30, -10, 20], [30, -10, 20], [30, -10, -10], [-30, -10, -10]]  # floor
    for i in range(1, nq):
        c = rng.uniform([-15, -6, -5], [15, 10, 12])
        u = rng.uniform(-6, 6, 3)
        v = np.cross(u, rng.uniform(-1, 1, 3))
        v = v / np.linalg.norm(v) * rng.uniform(2, 6)
        quads[i] = [c - u - v, c + u - v, c + u + v, c - u + v]
    spheres = np.zeros((ns, 4), dtype=np.float32)
    for i in range(ns):
        spheres[i] = list(rng.uniform([-16, -8, -2], [16, 6, 14])) + [rng.uniform(1.0, 3.0)]
    mats = np.zeros((nq + ns, 17), dtype=np.float32)
    for i in range(nq + ns):
        mats[i, 0:3] = rng.uniform(0.2, 0.9)
        if i == 1:
            mats[i, 3:6] = rng.uniform(5, 20, 3)
        mats[i, 6] = rng.choice([0.0, 0.02, 0.3])
        mats[i, 7] = rng.uniform(0, 0.5)
        mats[i, 8:11] = rng.uniform(0.5, 1.0, 3)
        mats[i, 11] = rng.uniform(1.05, 1.6)
        mats[i, 12] = rng.choice([0.0, 0.0, 1.0]) if i >= nq else 0.0
        mats[i, 13] = rng.uniform(0, 0.4)
        mats[i, 14:17] = rng.uniform(0, 1, 3)
    return quads, spheres, mats


def default_cornell_scene(simt_textured=False):
    """The literals of demofox_path_tracing_v2.cpp:320-454 (simt_textured.cpp:278-385) as arrays: (6, 4, 3) translated
    vertices, (3, 4) spheres, (9, 11) materials (albedo3, emissive3, specularColor3, percentSpecular, roughness)."""
    f = np.float32
    T = np.array([0.0, 0.0, 10.0], dtype=np.float32)
    quads = np.array([
        [[-12.6, -12.6, 25.0], [12.6, -12.6, 25.0], [12.6, 12.6, 25.0], [-12.6, 12.6, 25.0]],
        [[-12.6, -12.45, 25.0], [12.6, -12.45, 25.0], [12.6, -12.45, 15.0], [-12.6, -12.45, 15.0]],
        [[-12.6, 12.5, 25.0], [12.6, 12.5, 25.0], [12.6, 12.5, 15.0], [-12.6, 12.5, 15.0]],
        [[-12.5, -12.6, 25.0], [-12.5, -12.6, 15.0], [-12.5, 12.6, 15.0], [-12.5, 12.6, 25.0]],
        [[12.5, -12.6, 25.0], [12.5, -12.6, 15.0], [12.5, 12.6, 15.0], [12.5, 12.6, 25.0]],
        [[-5.0, 12.4, 22.5], [5.0, 12.4, 22.5], [5.0, 12.4, 17.5], [-5.0, 12.4, 17.5]]], dtype=np.float32) + T
    spheres = np.array([[x, -9.5, f(20.0) + f(10.0), f(3.0)] for x in (-9.0, 0.0, 9.0)], dtype=np.float32)
    m = np.zeros((9, 11), dtype=np.float32)
    m[0, 0:3] = m[1, 0:3] = m[2, 0:3] = 0.7
    m[3, 0:3] = [0.7, 0.1, 0.1]
    m[4, 0:3] = [0.1, 0.7, 0.1]
    m[5, 3:6] = np.array([1.0, 0.9, 0.7], dtype=np.float32) * f(20.0)
    if not simt_textured:
        m[6] = [0.9, 0.9, 0.5, 0, 0, 0, 0.9, 0.9, 0.9, 0.1, 0.2]
        m[7] = [0.9, 0.5, 0.9, 0, 0, 0, 0.9, 0.9, 0.9, 0.3, 0.2]
        m[8] = [0.0, 0.0, 1.0, 0, 0, 0, 1.0, 0.0, 0.0, 0.5, 0.4]
    else:
        m[6, 0:3] = [0.9, 0.9, 0.75]
        m[7, 0:3] = [0.9, 0.75, 0.9]
        m[8, 0:3] = [0.9, 0.75, 0.9]
    return quads, spheres, m


def random_cornell_scene(seed):
    """A scene of the Cornell family's shape (6 quads, 3 spheres, 9 materials) with other positions, sizes, orientations
    (quads no longer axis-aligned) and materials."""
    rng = np.random.default_rng(seed)
    quads, spheres, m = default_cornell_scene()
    for i in range(6):
        c = quads[i].mean(axis=0)
        u, v = quads[i][1] - quads[i][0], quads[i][3] - quads[i][0]
        n = np.cross(u, v)
        n /= np.linalg.norm(n)
        tilt = rng.uniform(-0.15, 0.15)
        u2 = u + n * tilt * np.linalg.norm(u)
        s = rng.uniform(0.7, 1.1)
        c2 = c + rng.uniform(-1.5, 1.5, 3)
        quads[i] = [c2 + s * (-u2 - v) / 2, c2 + s * (u2 - v) / 2, c2 + s * (u2 + v) / 2, c2 + s * (-u2 + v) / 2]
    for i in range(3):
        spheres[i, 0:3] += rng.uniform(-2.5, 2.5, 3)
        spheres[i, 3] = rng.uniform(1.5, 3.5)
    for i in range(9):
        m[i, 0:3] = rng.uniform(0.1, 0.9, 3)
        m[i, 6:9] = rng.uniform(0.3, 1.0, 3)
        m[i, 9] = rng.choice([0.0, 0.1, 0.5, 1.0])
        m[i, 10] = rng.uniform(0.0, 0.6)
    m[5, 3:6] = rng.uniform(5, 25, 3)
    m[2, 3:6] = rng.uniform(0, 1, 3)
    return quads.astype(np.float32), spheres.astype(np.float32), m.astype(np.float32)
