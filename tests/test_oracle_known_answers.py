"""Known-answer pins for the third-party arithmetic the oracle restates (no fixtures of Box2D / GL exist in
the reference): closed forms and invariants that any correct restatement must satisfy.

  * b2PolygonShape::ComputeMass / b2Body::ResetMassData: rectangle mass and rotational inertia in closed
    form, the hull as the sum of its four fixtures (parallel-axis theorem), evaluated in float64;
  * the GL point-sampling fill rule: a pixel-aligned rectangle covers exactly its pixels, two triangles
    sharing an edge cover every pixel of their union exactly once (watertight, no double hits), a
    sub-pixel sliver that contains no pixel centre draws nothing, clipping at the viewport border.
CPU only."""
import ctypes

import numpy as np
import pytest


def _fill(oracle, px, py, rgb=(1.0, 0.0, 0.0), img=None):
    L = oracle.lib()
    L.orc_raster_fill.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                  ctypes.c_float, ctypes.c_float, ctypes.c_float]
    if img is None:
        img = np.zeros((96, 96, 3), np.uint8)
    px = np.ascontiguousarray(px, np.float32); py = np.ascontiguousarray(py, np.float32)
    L.orc_raster_fill(img.ctypes.data, px.ctypes.data, py.ctypes.data, len(px), *[ctypes.c_float(c) for c in rgb])
    return img


def test_polygon_mass_closed_forms(oracle):
    w = oracle.OracleWorld(1)
    m = w.mass()        # hull mass, invMass, I, invI, lc.x, lc.y, wheel mass, invMass, I, invI, lc.x, lc.y
    SIZE = 0.02
    # wheel: 28 x 54 (x SIZE) rectangle, density 0.1, about its own centre
    ww, wh = 2 * 14 * SIZE, 2 * 27 * SIZE
    mass = 0.1 * ww * wh
    assert m[6] == pytest.approx(mass, rel=1e-6) and m[7] == pytest.approx(1 / mass, rel=1e-6)
    assert m[8] == pytest.approx(mass * (ww * ww + wh * wh) / 12.0, rel=1e-5)
    assert m[10] == 0.0 and m[11] == 0.0
    # hull: four convex fixtures, density 1 -- area / centroid / second moment by the shoelace formulas in float64
    polys = [w.shape(i).astype(np.float64) for i in range(4)]
    A = C = 0.0; Cxy = np.zeros(2); I0 = 0.0
    for P in polys:
        x, y = P[:, 0], P[:, 1]; xn, yn = np.roll(x, -1), np.roll(y, -1)
        cr = x * yn - xn * y
        a = cr.sum() / 2.0
        cx = ((x + xn) * cr).sum() / (6 * a); cy = ((y + yn) * cr).sum() / (6 * a)
        i_origin = (cr * (x * x + x * xn + xn * xn + y * y + y * yn + yn * yn)).sum() / 12.0
        A += a; Cxy += a * np.array([cx, cy]); I0 += i_origin
    Cxy /= A
    assert m[0] == pytest.approx(A, rel=1e-5)                         # density 1
    assert m[4] == pytest.approx(Cxy[0], abs=1e-6) and m[5] == pytest.approx(Cxy[1], rel=1e-4)
    assert m[2] == pytest.approx(I0 - A * (Cxy ** 2).sum(), rel=1e-4)  # b2Body stores I about the centre of mass


def test_fill_rule_known_answers(oracle):
    # pixel-aligned rectangle [10, 20) x [30, 35): exactly those pixels (row 0 of the image = top, y up in GL)
    img = _fill(oracle, [10, 20, 20, 10], [30, 30, 35, 35])
    cov = img[..., 0] > 0
    assert cov.sum() == 50 and cov[95 - 34:95 - 29, 10:20].all()
    # two triangles sharing the diagonal of an arbitrary quad: every pixel hit at most once, union == the quad's fill
    q = np.array([[5.3, 7.9], [71.2, 12.4], [80.6, 66.1], [11.7, 58.8]], np.float32)
    one = _fill(oracle, q[[0, 1, 2], 0], q[[0, 1, 2], 1], (1, 0, 0))[..., 0] > 0
    two = _fill(oracle, q[[0, 2, 3], 0], q[[0, 2, 3], 1], (1, 0, 0))[..., 0] > 0
    quad = _fill(oracle, q[:, 0], q[:, 1], (1, 0, 0))[..., 0] > 0
    assert not (one & two).any(), "shared edge drawn twice"
    assert np.array_equal(one | two, quad), "crack along the shared edge"
    area = 0.5 * abs(np.dot(q[:, 0], np.roll(q[:, 1], -1)) - np.dot(q[:, 1], np.roll(q[:, 0], -1)))
    assert abs(int(quad.sum()) - area) < 0.02 * area
    # a sliver that holds no pixel centre draws nothing; one that holds exactly one centre draws one pixel
    assert _fill(oracle, [40.6, 41.4, 41.4, 40.6], [50.6, 50.6, 51.4, 51.4]).sum() == 0
    assert (_fill(oracle, [40.4, 40.6, 40.6, 40.4], [50.4, 50.4, 50.6, 50.6])[..., 0] > 0).sum() == 1
    # clipping: a polygon larger than the viewport fills all 9216 pixels, one fully outside none
    assert (_fill(oracle, [-50, 200, 200, -50], [-50, -50, 200, 200])[..., 0] > 0).sum() == 96 * 96
    assert _fill(oracle, [100, 120, 120, 100], [10, 10, 20, 20]).sum() == 0
    # painter's order: the later polygon wins
    img = _fill(oracle, [0, 96, 96, 0], [0, 0, 96, 96], (0, 1, 0))
    img = _fill(oracle, [10, 20, 20, 10], [10, 10, 20, 20], (0, 0, 1), img)
    assert tuple(img[95 - 15, 15]) == (0, 0, 255) and tuple(img[95 - 50, 50]) == (0, 255, 0)


def test_revolute_joint_postconditions(oracle):
    """b2RevoluteJoint with limit + motor, as Box2D guarantees it after SolvePositionConstraints: every wheel
    stays pinned to its anchor on the hull (within a few b2_linearSlop) and inside the steering limits
    (+- 0.4 rad, within b2_angularSlop + one b2_maxAngularCorrection), under hard random steering / braking."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import action_tape, make_oracle_worlds
    tr, _ = oracle.generate_track(np.random.RandomState(17))
    w = make_oracle_worlds(oracle, [tr], [np.array([1, 0])], ['CCW'], 2, collisions=False)[0]
    w.step(None, render=False)
    tape = action_tape(17, 400, 1, 2, brake_p=0.5)
    SIZE = 0.02
    anchors = np.array([[-55, +80], [+55, +80], [-55, -82], [+55, -82]], np.float64) * SIZE
    worst_pin, worst_ang = 0.0, 0.0
    for s in range(400):
        w.step(tape[s, 0].astype(np.float64), render=False)
        b = w.bodies().astype(np.float64)                  # (A, 5, 9): p.x p.y angle v.x v.y w c.x c.y awake
        for c in range(2):
            hp, ha = b[c, 0, 0:2], b[c, 0, 2]
            R = np.array([[np.cos(ha), -np.sin(ha)], [np.sin(ha), np.cos(ha)]])
            for k in range(4):
                worst_pin = max(worst_pin, np.linalg.norm(b[c, 1 + k, 0:2] - (hp + R @ anchors[k])))
                worst_ang = max(worst_ang, abs(b[c, 1 + k, 2] - ha))
    assert worst_pin < 4 * 0.005, worst_pin
    assert worst_ang < 0.4 + 2.0 / 180 * np.pi + 8.0 / 180 * np.pi, worst_ang
    assert worst_ang > 0.3, "the tape must drive the front wheels to their limit"
