"""CPU restatement (numpy float32, operation for operation) of the list traversal kernel's per-round classification
(gpu_nbody_simulation_b200/csrc/traverse.cu: traverse_f32_list_kernel) and of its far-node force expression.

The kernel compares a node with the BOUNDING BOX of a warp's 64 bodies and skips the per-body test where the box
decides it for everybody (class A: all accept, class O: all open).  For each body to still see exactly the node set
of the reference's per-body walk (project.cu:731-772) these short cuts must never contradict the per-body
FP32 test `!(d2 <= thr)` — that is what the margins (kListDelta, the absolute slack) are for, and what this test checks
on random warps and nodes over the whole range of scales the scaled frame allows.  Second property: a FAR node
(distance from the box >= diag / 8 and >= 2^25 eps) evaluated with the single-float displacement and WITHOUT the
distance offset, G M / d^3, agrees with the reference's G M / (d^2 (d + eps)) in FP64 to a few FP32 ulps.

Test infrastructure only: nothing here is on the product path; no GPU needed."""
import numpy as np

F = np.float32
K_DELTA = F(1e-4)        # kListDelta
ULP = F(1.2e-7)


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def warp_frame(sx, sy, feps):
    """prologue of the kernel for one warp: sx, sy = scaled FP64 coordinates of its 64 bodies"""
    fx, fy = sx.astype(F), sy.astype(F)
    x0, x1, y0, y1 = fx.min(), fx.max(), fy.min(), fy.max()
    ox = 0.5 * (np.float64(x0) + np.float64(x1)); oy = 0.5 * (np.float64(y0) + np.float64(y1))
    rx, ry = sx - ox, sy - oy
    xh, yh = rx.astype(F), ry.astype(F)
    w = dict(nxh=-xh, nyh=-yh, nxl=-(rx - xh.astype(np.float64)).astype(F), nyl=-(ry - yh.astype(np.float64)).astype(F))
    w["oxh"] = F(ox); w["oxl"] = F(ox - np.float64(w["oxh"])); w["oyh"] = F(oy); w["oyl"] = F(oy - np.float64(w["oyh"]))
    mag = max(abs(x0), abs(x1), abs(y0), abs(y1))
    w["bx0"] = F(np.float64(x0) - ox); w["bx1"] = F(np.float64(x1) - ox)
    w["by0"] = F(np.float64(y0) - oy); w["by1"] = F(np.float64(y1) - oy)
    w["slack"] = F(ULP * mag)
    wx, wy = F(w["bx1"] - w["bx0"]), F(w["by1"] - w["by0"])
    far2 = F(F(0.015625) * F(F(wx * wx) + F(wy * wy)))
    dfar = F(F(33554432.0) * feps)
    w["far2"] = max(far2, F(dfar * dfar))
    return w


def to_local(rh, rl, oh, ol):
    a, b = rh, F(-oh)
    s = F(a + b); bb = F(s - a)
    e = F(F(a - F(s - bb)) + F(b - bb))
    t = F(F(rl - ol) + e)
    hi = F(s + t)
    return hi, F(t - F(hi - s))


def classify(w, cx, cy, thr):
    """one internal node (thr >= 0, non-zero mass) against the warp's box; returns (class, local double-float COM)"""
    chx, chy = F(cx), F(cy)
    clx, cly = F(cx - np.float64(chx)), F(cy - np.float64(chy))
    lx, lzx = to_local(chx, clx, w["oxh"], w["oxl"]); ly, lzy = to_local(chy, cly, w["oyh"], w["oyl"])
    ex = F(w["slack"] + F(ULP * abs(lx))); ey = F(w["slack"] + F(ULP * abs(ly)))
    nx = max(F(max(F(w["bx0"] - lx), F(lx - w["bx1"])) - ex), F(0)); ny = max(F(max(F(w["by0"] - ly), F(ly - w["by1"])) - ey), F(0))
    dmin2 = F(F(nx * nx) + F(ny * ny))
    fx = F(max(F(lx - w["bx0"]), F(w["bx1"] - lx)) + ex); fy = F(max(F(ly - w["by0"]), F(w["by1"] - ly)) + ey)
    dmax2 = F(F(fx * fx) + F(fy * fy))
    if dmin2 > F(thr * F(F(1) + K_DELTA)):
        cls = "A_near" if not (dmin2 >= w["far2"]) else "A_far"
    elif dmax2 <= F(thr * F(F(1) - K_DELTA)):
        cls = "O"
    else:
        cls = "M"
    return cls, (lx, ly, lzx, lzy)


def per_body_d2(w, L):
    """eval_mixed: double-float displacement, packed FP32 arithmetic"""
    lx, ly, lzx, lzy = L
    dx = (F(lx) + w["nxh"]) + (F(lzx) + w["nxl"])
    dy = (F(ly) + w["nyh"]) + (F(lzy) + w["nyl"])
    return fma32(dx, dx, (dy * dy).astype(F)), dx, dy


def random_case(rng):
    """a warp of 64 bodies somewhere in the scaled frame ([-2^21, 2^21]) and a node at a random distance from it"""
    centre = rng.uniform(-2.0 ** 20.5, 2.0 ** 20.5, 2)
    spread = 2.0 ** rng.uniform(-12, 17)                         # from coincident-ish clusters to a 16th of the box
    sx = centre[0] + spread * rng.uniform(-1, 1, 64); sy = centre[1] + spread * rng.uniform(-1, 1, 64)
    if rng.random() < 0.3:                                       # just outside the box: near class-A candidates
        gap = spread * 2.0 ** rng.uniform(-8, -2)
        cx, cy = centre[0] + spread + gap, centre[1] + spread * rng.uniform(-1, 1)
        thr = F((gap * 2.0 ** rng.uniform(-3, 0.5)) ** 2)
    else:
        dist = spread * 2.0 ** rng.uniform(-4, 7)
        ang = rng.uniform(0, 2 * np.pi)
        cx, cy = centre[0] + dist * np.cos(ang), centre[1] + dist * np.sin(ang)
        thr = F((dist * 2.0 ** rng.uniform(-1.5, 1.5)) ** 2)     # acceptance threshold around the actual distance
    return sx, sy, cx, cy, thr


def test_box_shortcuts_never_contradict_the_per_body_test():
    rng = np.random.default_rng(20261018)
    feps = F(1e-15 * 2.0 ** 22)                                   # dist_eps x a typical power-of-two scale
    seen = {"A_far": 0, "A_near": 0, "O": 0, "M": 0}
    for _ in range(6000):
        sx, sy, cx, cy, thr = random_case(rng)
        w = warp_frame(sx, sy, feps)
        cls, L = classify(w, cx, cy, thr)
        d2, _, _ = per_body_d2(w, L)
        accept = ~(d2 <= thr)                                     # the per-body test of the pair kernel / class M
        seen[cls] += 1
        if cls.startswith("A"):
            assert accept.all(), (cls, float(thr), d2.min())
        elif cls == "O":
            assert (~accept).all(), (cls, float(thr), d2.max())
    assert min(seen.values()) > 150, seen                         # every class is exercised


def test_far_nodes_single_float_displacement_without_the_offset():
    rng = np.random.default_rng(7)
    eps = 1e-15 * 2.0 ** 22
    worst = 0.0
    n_far = 0
    for _ in range(6000):
        sx, sy, cx, cy, thr = random_case(rng)
        w = warp_frame(sx, sy, F(eps))
        cls, L = classify(w, cx, cy, F(0))                        # thr = 0: everything outside the box is class A
        if cls != "A_far":
            continue
        n_far += 1
        gm = F(3.0)
        dx = F(L[0]) + w["nxh"]; dy = F(L[1]) + w["nyh"]           # apply_far: ONE add per coordinate
        d2 = fma32(dx, dx, (dy * dy).astype(F))
        inv = (1.0 / np.sqrt(d2.astype(np.float64))).astype(F)    # MUFU.RSQ (the approximation's 2^-22 is not modelled)
        g = (F(gm) * inv).astype(F) * (inv * inv).astype(F)
        fx32 = (g * dx).astype(np.float64)
        ddx, ddy = cx - sx, cy - sy                               # reference expression in FP64, project.cu:765-769
        d = np.sqrt(ddx * ddx + ddy * ddy)
        fx64 = 3.0 / (d * d * (d + eps)) * ddx
        f64 = 3.0 / (d * d * (d + eps)) * d
        worst = max(worst, float(np.max(np.abs(fx32 - fx64) / f64)))
    assert n_far > 1500
    assert worst <= 4e-6, worst                                   # 2^-24 (2 + 2 diag / d) <= 1.1e-6 on d, x3 on d^-3 + roundings
