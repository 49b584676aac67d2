"""Seeded synthetic inputs for the deformable-aggregation tests (shared by CPU and GPU suites)."""
import numpy as np

# (H, W) of the 4 FPN levels (strides 4/8/16/32) for the input resolutions in BASELINE.json
LEVELS_352x640 = [(88, 160), (44, 80), (22, 40), (11, 20)]
LEVELS_256x704 = [(64, 176), (32, 88), (16, 44), (8, 22)]
LEVELS_512x1408 = [(128, 352), (64, 176), (32, 88), (16, 44)]


def level_tables(level_hw, cams):
    shapes = np.array([level_hw] * cams, dtype=np.int32)                      # [cams, L, 2]
    sizes = (shapes[..., 0] * shapes[..., 1]).reshape(-1)
    starts = (np.cumsum(sizes) - sizes).reshape(cams, -1).astype(np.int32)    # absolute rows
    return shapes, starts, int(sizes.sum())


def random_locations(rng, shape, frac_inside=0.35, band=True):
    """(x,y): a mix of strictly-inside, border-hugging, exactly-0/1, just-outside and far-outside."""
    n = int(np.prod(shape))
    xy = rng.uniform(-0.6, 1.6, size=(n, 2))
    inside = rng.random(n) < frac_inside
    xy[inside] = rng.uniform(0.0, 1.0, size=(int(inside.sum()), 2))
    hug = rng.random(n) < 0.08
    xy[hug] = np.where(rng.random((int(hug.sum()), 2)) < 0.5,
                       rng.uniform(0, 0.01, (int(hug.sum()), 2)), rng.uniform(0.99, 1.0, (int(hug.sum()), 2)))
    if band:
        exact = rng.random(n) < 0.02
        xy[exact] = rng.choice([0.0, 1.0, 0.5], size=(int(exact.sum()), 2))
    return xy.reshape(*shape, 2).astype(np.float32)


def make_case(seed, bs, cams, level_hw, C, G, A, P, frac_inside=0.35, weight_kind="softmax"):
    rng = np.random.default_rng(seed)
    shapes, starts, F = level_tables(level_hw, cams)
    L = len(level_hw)
    feat = rng.standard_normal((bs, F, C), dtype=np.float32)
    loc = random_locations(rng, (bs, A, P, cams), frac_inside)
    logits = rng.standard_normal((bs, A, P, cams, L, G), dtype=np.float32)
    if weight_kind == "softmax":
        m = logits.reshape(bs, A, -1, G)
        e = np.exp(m - m.max(axis=2, keepdims=True))
        w = (e / e.sum(axis=2, keepdims=True)).reshape(logits.shape).astype(np.float32)
    else:
        w = logits
    grad_out = rng.standard_normal((bs, A, C), dtype=np.float32)
    return dict(feat=feat, shapes=shapes, starts=starts, loc=loc, weights=w, grad_out=grad_out,
                dims=(bs, cams, F, C, L, A, P, G))


# ---- realistic geometry: Bench2Drive cameras + box anchors (constants restated from
# bench2drive/leaderboard/team_code/hipad_b2d_agent.py:39-69 of the reference; data, not code)
LIDAR2IMG = np.array([
    [[1142.51841, 800.0, 0.0, -952.0], [0.0, 450.0, -1142.51841, -809.704417], [0.0, 1.0, 0.0, -1.19], [0, 0, 0, 1]],
    [[1310.64327, -477.035138, 0.0, -406.010608], [368.61842, 258.109396, -1142.51841, -647.29675],
     [0.819152044, 0.573576436, 0.0, -0.829094072], [0, 0, 0, 1]],
    [[6.03961325e-14, 1394.75744, 0.0, -920.539908], [-368.61842, 258.109396, -1142.51841, -647.29675],
     [-0.819152044, 0.573576436, 0.0, -0.829094072], [0, 0, 0, 1]],
    [[-560.166031, -800.0, 0.0, -1288.0], [5.5109106e-14, -450.0, -560.166031, -858.939847],
     [1.2246468e-16, -1.0, 0.0, -1.61], [0, 0, 0, 1]],
    [[-1142.51841, 800.0, 0.0, -684.385123], [-422.861679, -153.909064, -1142.51841, -496.004706],
     [-0.939692621, -0.342020143, 0.0, -0.492889531], [0, 0, 0, 1]],
    [[360.989788, -1347.23223, 0.0, -104.238127], [422.861679, -153.909064, -1142.51841, -496.004706],
     [0.939692621, -0.342020143, 0.0, -0.492889531], [0, 0, 0, 1]],
], dtype=np.float64)


def projection_matrices(final_hw, src_hw=(900, 1600)):
    """aug (resize + top crop, hipad_b2d_agent.py:421-443) @ LIDAR2IMG -> [6,4,4] float32."""
    fH, fW = final_hw
    H, W = src_hw
    resize = max(fH / H, fW / W)
    newW, newH = int(W * resize), int(H * resize)
    aug = np.eye(4)
    aug[0, 0] = aug[1, 1] = resize
    aug[0, 3] = -int(max(0, newW - fW) / 2)
    aug[1, 3] = -(newH - fH)
    return (aug @ LIDAR2IMG).astype(np.float32)


def synthetic_anchor_points(rng, kind, A, P):
    """3-D key points [A,P,3] with the spatial statistics of the shipped k-means anchors:
    det boxes scattered up to ~50 m around ego, map poly-lines on the ground plane, plan
    waypoints fanning out ahead of ego (the .npy files themselves live in the reference)."""
    if kind == "det":
        r = rng.uniform(2.0, 55.0, A)
        th = rng.uniform(-np.pi, np.pi, A)
        ctr = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(-0.9, 0.4, A)], 1)
        size = np.exp(rng.normal([0.7, 1.5, 0.5], 0.25, (A, 3)))
        off = rng.uniform(-0.5, 0.5, (A, P, 3)) * size[:, None]
        return (ctr[:, None] + off).astype(np.float32)
    if kind == "map":
        n_s = max(1, P // 15)
        x0 = rng.uniform(-15, 15, A)
        y0 = rng.uniform(-30, 30, A)
        ang = rng.uniform(-np.pi, np.pi, A)
        t = np.linspace(0, 30, n_s)
        line = np.stack([x0[:, None] + t * np.cos(ang)[:, None], y0[:, None] + t * np.sin(ang)[:, None]], -1)
    else:  # plan: waypoints ahead (+y is forward in the B2D lidar frame)
        n_s = max(1, P // 15)
        speed = rng.uniform(0.2, 5.0, A)
        curve = rng.normal(0, 0.04, A)
        t = np.arange(1, n_s + 1) * 1.0
        yy = speed[:, None] * t
        line = np.stack([curve[:, None] * yy ** 2, yy], -1)
    per = P // n_s
    pts = np.repeat(line, per, axis=1) + rng.normal(0, 0.5, (A, n_s * per, 2))
    heights = np.tile(np.repeat(np.array([0, 0.5, -0.5, 1, -1.0]), max(1, per // 5)), n_s)[: n_s * per]
    z = -1.84023 + heights[None].repeat(A, 0)
    out = np.concatenate([pts, z[..., None]], -1)
    if out.shape[1] < P:
        out = np.concatenate([out, out[:, : P - out.shape[1]]], 1)
    return out.astype(np.float32)


def project(points, proj, image_wh):
    """numpy float32 restatement of project_points for input generation: [A,P,3] -> [A,P,cams,2]."""
    hom = np.concatenate([points, np.ones_like(points[..., :1])], -1).astype(np.float32)
    cam = np.einsum("cij,apj->apci", proj.astype(np.float32), hom)
    xy = cam[..., :2] / np.maximum(cam[..., 2:3], np.float32(1e-5))
    return (xy / np.asarray(image_wh, np.float32)).astype(np.float32)


def make_geo_case(seed, kind, bs, level_hw, final_hw, C=256, G=8, A=900, P=13, with_feat=True):
    """Stage-2-shaped call with realistically distributed sampling locations (10-20 % visible)."""
    rng = np.random.default_rng(seed)
    cams = 6
    shapes, starts, F = level_tables(level_hw, cams)
    L = len(level_hw)
    proj = projection_matrices(final_hw)
    kp = np.stack([synthetic_anchor_points(rng, kind, A, P) for _ in range(bs)])
    loc = np.stack([project(kp[b], proj, (final_hw[1], final_hw[0])) for b in range(bs)])
    logits = rng.standard_normal((bs, A, cams, L, P, G), dtype=np.float32)
    m = logits.transpose(0, 1, 5, 2, 3, 4).reshape(bs, A, G, -1)
    e = np.exp(m - m.max(-1, keepdims=True))
    w = (e / e.sum(-1, keepdims=True)).reshape(bs, A, G, cams, L, P).transpose(0, 1, 5, 3, 4, 2)
    case = dict(shapes=shapes, starts=starts, loc=loc.astype(np.float32),
                weights=np.ascontiguousarray(w, dtype=np.float32), key_points=kp, logits=logits,
                projection_mat=np.broadcast_to(proj, (bs, cams, 4, 4)).copy(),
                image_wh=np.broadcast_to(np.array([final_hw[1], final_hw[0]], np.float32), (bs, cams, 2)).copy(),
                dims=(bs, cams, F, C, L, A, P, G))
    if with_feat:
        case["feat"] = rng.standard_normal((bs, F, C), dtype=np.float32)
        case["grad_out"] = rng.standard_normal((bs, A, C), dtype=np.float32)
    return case
