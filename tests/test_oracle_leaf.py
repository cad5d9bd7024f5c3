"""CPU: the oracle's leaf predicates against independent float64 restatements and the sanity vectors of
SURVEY.md Appendix E (the only pins that exist: the reference has no tests for this path)."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from orbit_b200 import scenes


def test_frustum_planes_appendix_e():
    v = scenes.perspective_view((0, 0, 0), (0, 0, -1), 1920, 1080)
    expect = np.array([[0.4903, 0, -0.8716, 0], [-0.4903, 0, -0.8716, 0], [0, 0.7071, -0.7071, 0], [0, -0.7071, -0.7071, 0], [0, 0, -1, 0.01]])
    assert v.planes.shape == (5, 4)
    assert np.allclose(v.planes, expect, atol=1e-4)


def test_project_sphere_appendix_e(oracle):
    lib = oracle.lib()
    lib.oracle_project_sphere.argtypes = [C.c_float] * 8 + [C.POINTER(C.c_float)]
    out = (C.c_float * 6)()
    lib.oracle_project_sphere(1.0, 0.5, -5.0, 0.5, 1.0, 0.01, 0.5625, 1.0, out)
    assert np.allclose(list(out)[:4], [0.5280, 0.3990, 0.5857, 0.5000], atol=1e-4)
    assert abs(out[4] - 0.0022222) < 1e-7 and out[5] == 1.0
    # brute force: project 2e5 surface points
    rng = np.random.default_rng(1)
    p = rng.normal(size=(200000, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    p = p * 0.5 + np.array([1.0, 0.5, -5.0])
    x = 0.5 + 0.5 * (0.5625 * p[:, 0] / -p[:, 2]); y = 0.5 - 0.5 * (1.0 * p[:, 1] / -p[:, 2])
    assert abs(x.min() - out[0]) < 2e-4 and abs(x.max() - out[2]) < 2e-4 and abs(y.min() - out[1]) < 2e-4 and abs(y.max() - out[3]) < 2e-4
    # on a 1024^2 pyramid: 59.0 x 103.4 px -> lambda 6.69 -> level 7
    assert lib.oracle_hiz_level(C.c_float(max((out[2] - out[0]) * 1024, (out[3] - out[1]) * 1024)), 11) == 7


def test_mip_level_rule_appendix_e(oracle):
    lib = oracle.lib()
    table = {0.3: 0, 0.9: 0, 1.0: 0, 1.41: 0, 1.42: 1, 2.8: 1, 2.9: 2, 5.6: 2, 5.7: 3, 1000.0: 10, 5000.0: 10, 0.0: 0, -3.0: 0,
             float("inf"): 10, float("nan"): 0}
    for x, lvl in table.items():
        assert lib.oracle_hiz_level(C.c_float(x), 11) == lvl, x


@settings(max_examples=300, deadline=None)
@given(st.floats(min_value=2.0 ** -100, max_value=2.0 ** 100, allow_nan=False, width=32))
def test_mip_level_equals_rounded_log2(x):
    import oracle_ref
    lvl = oracle_ref.lib().oracle_hiz_level(C.c_float(x), 16)
    lam = np.log2(np.float64(np.float32(x)))
    want = int(np.clip(np.ceil(lam + 0.5) - 1, 0, 15))
    if abs((lam + 0.5) - round(lam + 0.5)) > 1e-6:   # away from exact ties (which are irrational in x)
        assert lvl == want


def test_log2_contract_function(oracle):
    rng = np.random.default_rng(7)
    xs = np.concatenate([np.exp(rng.uniform(-80, 80, 4000)), rng.uniform(0.5, 2.0, 4000), [1.0, 2.0, 0.5, 1e-40, 3e-39, 16.0]]).astype(np.float32)
    got = np.array([oracle.log2f(float(x)) for x in xs])
    want = np.log2(xs.astype(np.float64))
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
    assert err.max() < 4e-7, err.max()     # a few ulp
    assert oracle.log2f(1.0) == 0.0 and oracle.log2f(8.0) == 3.0 and oracle.log2f(0.25) == -2.0
    assert oracle.log2f(0.0) == -np.inf and np.isnan(oracle.log2f(-1.0)) and oracle.log2f(np.inf) == np.inf


def _numpy_reduce_min(src, dw, dh):
    """Independent restatement of the ReduceMin footprint rule in numpy, level by level."""
    sh, sw = src.shape
    xs = (np.arange(dw, dtype=np.float32) + np.float32(0.5)) / np.float32(dw)
    ys = (np.arange(dh, dtype=np.float32) + np.float32(0.5)) / np.float32(dh)
    fx = np.floor(xs * np.float32(sw) - np.float32(0.5)).astype(np.int64)
    fy = np.floor(ys * np.float32(sh) - np.float32(0.5)).astype(np.int64)
    x0, x1 = np.clip(fx, 0, sw - 1), np.clip(fx + 1, 0, sw - 1)
    y0, y1 = np.clip(fy, 0, sh - 1), np.clip(fy + 1, 0, sh - 1)
    return np.minimum(np.minimum(src[np.ix_(y0, x0)], src[np.ix_(y0, x1)]), np.minimum(src[np.ix_(y1, x0)], src[np.ix_(y1, x1)]))


@pytest.mark.parametrize("size", [(1920, 1080), (100, 60), (64, 64), (129, 257), (3, 2), (4096, 33)])
def test_hiz_build_against_numpy_restatement(oracle, size):
    w, h = size
    rng = np.random.default_rng(w + 31 * h)
    depth = rng.random((h, w), dtype=np.float32)
    depth[rng.random((h, w)) < 0.25] = 0.0
    info, texels = oracle.hiz_build(depth)
    src = depth
    for l in range(info.levels):
        lw, lh = max(info.width >> l, 1), max(info.height >> l, 1)
        want = _numpy_reduce_min(src, lw, lh)
        got = texels[info.level_offset[l]:info.level_offset[l] + lw * lh].reshape(lh, lw)
        assert np.array_equal(got, want), ("level", l)
        src = want
    if (w, h) == (1920, 1080):
        # Appendix B: 1920 -> 1024 footprint in x is floor(1.875 x + 0.4375) and does NOT cover column 1 for x = 1
        depth2 = np.ones((h, w), np.float32); depth2[:, 1] = 0.25
        _, t2 = oracle.hiz_build(depth2)
        row0 = t2[:1024]
        assert row0[0] == 0.25 and row0[1] == 1.0


def test_cone_cull_matches_float64(oracle):
    lib = oracle.lib()
    lib.oracle_cone_cull.argtypes = [C.c_float] * 8
    rng = np.random.default_rng(3)
    n_checked = 0
    for _ in range(4000):
        c = rng.normal(size=3) * 10; a = rng.normal(size=3); a /= np.linalg.norm(a)
        r, cut = rng.uniform(0.1, 1.0), rng.uniform(-1, 1)
        lhs, rhs = float(np.dot(c, a)), float(cut * np.linalg.norm(c) + r)
        if abs(lhs - rhs) < 1e-4 * max(abs(lhs), abs(rhs), 1.0):
            continue
        n_checked += 1
        assert bool(lib.oracle_cone_cull(*[C.c_float(v) for v in (*c, r, *a, cut)])) == (lhs >= rhs)
    assert n_checked > 3000


def _rust_twin_project_sphere_clip_space(c, r, znear, p00, p11):
    """numpy float32 restatement of the reference's OWN CPU twin of the shader math, math::project_sphere_clip_space
    (src/math.rs:170-199; Rust f32 arithmetic is unfused). Returns None or the clip-space aabb (minx, miny, maxx, maxy)."""
    f = np.float32
    cx_, cy_, cz_ = f(c[0]), f(c[1]), f(c[2]); r = f(r)
    if cz_ < f(r + f(znear)):
        return None
    def axis(a):
        c0, c1 = f(-a), f(-cz_)
        v0 = f(np.sqrt(f(f(f(c0 * c0) + f(c1 * c1)) - f(r * r)))); v1 = r
        mn = (f(f(v0 * c0) + f(f(-v1) * c1)), f(f(v1 * c0) + f(v0 * c1)))
        mx = (f(f(v0 * c0) + f(v1 * c1)), f(f(f(-v1) * c0) + f(v0 * c1)))
        return mn, mx
    (minx, maxx), (miny, maxy) = axis(cx_), axis(cy_)
    return (f(f(minx[0] / minx[1]) * f(p00)), f(f(miny[0] / miny[1]) * f(p11)), f(f(maxx[0] / maxx[1]) * f(p00)), f(f(maxy[0] / maxy[1]) * f(p11)))


def test_project_sphere_against_reference_cpu_twin(oracle):
    """The only independent restatement of the projection inside the reference: agree within the north star's 1e-5
    (the shader path has Fma where the Rust twin has separate ops, so not bit-for-bit)."""
    lib = oracle.lib()
    lib.oracle_project_sphere.argtypes = [C.c_float] * 8 + [C.POINTER(C.c_float)]
    rng = np.random.default_rng(21)
    out = (C.c_float * 6)()
    n_some = n_none = 0
    for _ in range(3000):
        c = rng.uniform([-30, -30, 0.05], [30, 30, 200]); r = rng.uniform(0.05, 3.0)
        twin = _rust_twin_project_sphere_clip_space(c, r, 0.01, 0.5625, 1.0)
        # the shader negates view-space z before calling project_sphere: pass z_view = -c.z
        lib.oracle_project_sphere(C.c_float(c[0]), C.c_float(c[1]), C.c_float(-c[2]), C.c_float(r), C.c_float(1.0), C.c_float(0.01),
                                  C.c_float(0.5625), C.c_float(1.0), out)
        margin = abs(c[2] - (r + 0.01))
        if margin > 1e-4 * max(c[2], 1.0):
            assert (twin is not None) == (out[5] == 1.0)
        if twin is None:
            n_none += 1
            continue
        n_some += 1
        uv = np.array([twin[0] * 0.5 + 0.5, twin[3] * -0.5 + 0.5, twin[2] * 0.5 + 0.5, twin[1] * -0.5 + 0.5], np.float64)
        got = np.array(list(out)[:4], np.float64)
        assert np.allclose(got, uv, rtol=1e-5, atol=1e-5), (c, r, got, uv)
    assert n_some > 1000 and n_none > 10


def test_orthographic_box_against_shadow_renderer_twin(oracle):
    """shadow_renderer.rs:593-604 restates the ortho screen box + depth on the CPU (with width AND height recip,
    where the shader uses p00 for both — entity_cull.comp:166; equal for the square cascades the reference renders)."""
    from orbit_b200 import layouts as L
    import oracle_ref
    f = np.float32
    rng = np.random.default_rng(5)
    sc, _ = scenes.config_c1(scale=0.1)
    view = scenes.orthographic_view((0, 50, 0), (0.3, -1.0, 0.2), 2048, 2048, half_width=40.0, near=-20.0, far=150.0)
    g = oracle_ref.gpu_cull_info(view, "write")
    wr2 = f(g.p00_or_width_recip_x2)
    assert abs(wr2 - 2.0 / 80.0) < 1e-7 and g.p11_or_height_recip_x2 == g.p00_or_width_recip_x2
    for _ in range(200):
        cx, cy, cz, r = rng.uniform(-60, 60), rng.uniform(-60, 60), rng.uniform(-140, 10), rng.uniform(0.1, 5)
        centre = (f(wr2 * f(cx)), f(wr2 * f(cy))); box = f(f(r) * wr2)
        aabb = np.clip([centre[0] - box, centre[1] - box, centre[0] + box, centre[1] + box], -1, 1)
        rr = f(1.0) / f(f(g.z_far) - f(g.z_near))
        depth_twin = f(f(f(cz) + f(r)) * rr) + f(rr * f(g.z_far))
        # oracle formula (SPIR-V): k * (fma(r_model, s, z) + z_far), s = 1
        depth_oracle = rr * f(f(np.float64(r) * 1.0 + np.float64(cz)) + f(g.z_far))
        assert abs(float(depth_twin) - float(depth_oracle)) <= 1e-5 * max(abs(float(depth_twin)), 1e-3)
        assert np.all(aabb >= -1) and np.all(aabb <= 1)
