"""Stub third-party modules (Box2D, gym, pyglet, shapely) for running the UNMODIFIED reference
module /root/reference/gym_multi_car_racing/multi_car_racing.py in a container where none of
them is installable.

The stubs implement only the surface the reference touches.  Their arithmetic is the CPU
oracle's (oracle/mcr_oracle.c through ctypes): rigid bodies / tyre model / overlap predicate /
polygon fill.  What this buys: every line of the reference's OWN Python -- FrictionDetector,
reset() RNG order + spawn grid, _create_track, step() reward/backward/done block, the camera
maths of _render_window, the draw lists and order of render_road / Car.draw hand-off /
render_indicators, the read-back flip -- executes for real, and its outputs become golden
fixtures that pin the oracle's restatement of those lines.  The third-party arithmetic itself
stays "parity unpinned" (see oracle/mcr_oracle.c).

Used only by tests/golden/make_golden.py (run in the build container, not on the GPU box).
"""
import ctypes
import math
import sys
import types

import numpy as np

import mcr_oracle as mo

F32 = np.float32
_state = {"window": None}


# ------------------------------------------------------------------------------------------
# Box2D
# ------------------------------------------------------------------------------------------
class _Vec2(tuple):
    x = property(lambda s: s[0])
    y = property(lambda s: s[1])


class polygonShape:
    def __init__(self, vertices=None, **kw):
        self.vertices = list(vertices) if vertices is not None else []


class fixtureDef:
    def __init__(self, shape=None, **kw):
        self.shape = shape
        self.__dict__.update(kw)


class _Fixture:
    def __init__(self, body, vertices):
        self.body = body
        self.shape = polygonShape(vertices)
        self.sensor = False


class _TileBody:
    def __init__(self, world, vertices, index):
        self.world, self.index = world, index
        self.userData = None
        self.fixtures = [_Fixture(self, [tuple(v) for v in vertices])]


class contactListener:
    def __init__(self):
        pass


class _Contact:
    def __init__(self, fa, fb):
        self.fixtureA, self.fixtureB = fa, fb


class edgeShape:  # imported by the reference, unused
    pass


class circleShape:
    pass


class revoluteJointDef:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class b2World:
    """One oracle world per episode (built lazily at the first Step after the reference created
    tiles and cars -- a fresh b2World's contact order, D1)."""

    def __init__(self, gravity=(0, 0), contactListener=None):
        self.listener = contactListener
        self.tiles, self.cars = [], []
        self.ow = None
        self.prev_touch = set()

    def CreateStaticBody(self, fixtures=None, **kw):
        t = _TileBody(self, fixtures.shape.vertices, len(self.tiles))
        self.tiles.append(t)
        self.ow = None
        return t

    def DestroyBody(self, body):
        if body in self.tiles:
            self.tiles.remove(body)
        self.ow = None
        self.prev_touch = set()

    def _register_car(self, car):
        self.cars.append(car)
        self.ow = None

    def _unregister_car(self, car):
        if car in self.cars:
            self.cars.remove(car)
        self.ow = None

    def _ensure(self):
        if self.ow is not None:
            return
        A = len(self.cars)
        for i, t in enumerate(self.tiles):
            t.index = i
        quads = np.array([t.fixtures[0].shape.vertices for t in self.tiles], dtype=np.float64).reshape(-1, 4, 2)
        T = len(quads)
        nodes = np.zeros((T, 4))
        tr = mo.Track(nodes, quads, np.full((T, 3), 0.4, np.float32), np.arange(T, dtype=np.int32), (0, 0))
        self.ow = mo.OracleWorld(A)
        self.ow.set_track(tr, False)
        self.ow.spawn(np.array([[c.init_angle, c.init_x, c.init_y] for c in self.cars], dtype=np.float64))
        for i, c in enumerate(self.cars):
            c.index = i
        self.prev_touch = set()
        L = self.ow.L
        L.orc_ext_steer.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        L.orc_ext_gas.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        L.orc_ext_brake.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        L.orc_ext_car_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
        L.orc_ext_collide_pairs.restype = ctypes.c_int
        L.orc_ext_collide_pairs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_ext_solve.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_int]

    def Step(self, dt, vel_iters, pos_iters):
        self._ensure()
        L, h = self.ow.L, self.ow.h
        buf = np.zeros((4096, 4), np.int32)
        n = L.orc_ext_collide_pairs(h, buf.ctypes.data, 4096)
        now = [tuple(int(v) for v in buf[i]) for i in range(n)]          # (tile, car, fixture, active)
        now_keys = {p[:3] for p in now}
        active = {p[:3]: p[3] for p in now}
        # b2Contact::Update: transitions fire the listener in contact-list order (tile desc, car desc, fixture desc)
        order = sorted(now_keys | self.prev_touch, key=lambda k: (-k[0], -k[1], -k[2]))
        new_prev = set()
        for key in order:
            t, c, f = key
            car = self.cars[c]
            body = car.wheels[f] if f < 4 else car.hull
            was, is_now = key in self.prev_touch, key in now_keys
            is_active = active.get(key, body._awake())
            if not is_active:                       # a sleeping body's contacts are not updated
                if was:
                    new_prev.add(key)
                continue
            if is_now:
                new_prev.add(key)
            if is_now and not was:
                self.listener.BeginContact(_Contact(self.tiles[t].fixtures[0], body._fixture(f)))
            elif was and not is_now:
                self.listener.EndContact(_Contact(self.tiles[t].fixtures[0], body._fixture(f)))
        self.prev_touch = new_prev
        L.orc_ext_solve(h, float(dt), int(vel_iters), int(pos_iters))


# ------------------------------------------------------------------------------------------
# gym.envs.box2d.car_dynamics
# ------------------------------------------------------------------------------------------
SIZE = 0.02
WHEEL_R = 27
WHEEL_W = 14
WHEELPOS = [(-55, +80), (+55, +80), (-55, -82), (+55, -82)]
WHEEL_COLOR = (0.0, 0.0, 0.0)
WHEEL_WHITE = (0.3, 0.3, 0.3)


class _Transform:
    """b2Transform * v in fp32 (b2Mul: (q.c*x - q.s*y) + p.x)."""

    def __init__(self, px, py, qs, qc):
        self.px, self.py, self.qs, self.qc = F32(px), F32(py), F32(qs), F32(qc)

    def __mul__(self, v):
        x, y = F32(v[0]), F32(v[1])
        return (F32(F32(self.qc * x) - F32(self.qs * y)) + self.px, F32(F32(self.qs * x) + F32(self.qc * y)) + self.py)


class _Body:
    def __init__(self, car, bi):
        self.car, self.bi = car, bi

    def _row(self):
        return self.car.world.ow.bodies()[self.car.index, self.bi]

    def _awake(self):
        return bool(self._row()[8])

    position = property(lambda s: _Vec2((float(s._row()[0]), float(s._row()[1]))))
    angle = property(lambda s: float(s._row()[2]))
    linearVelocity = property(lambda s: _Vec2((float(s._row()[3]), float(s._row()[4]))))
    angularVelocity = property(lambda s: float(s._row()[5]))

    @property
    def transform(self):
        r = self._row()
        a = F32(r[2])
        return _Transform(r[0], r[1], F32(math.sin(float(a))), F32(math.cos(float(a))))


class _ShapeView:
    def __init__(self, verts):
        self.vertices = verts


class _FixtureView:
    def __init__(self, body, verts):
        self.body, self.shape = body, _ShapeView(verts)


class _Joint:
    def __init__(self, wheel):
        self.wheel = wheel

    @property
    def angle(self):
        b = self.wheel.car.world.ow.bodies()[self.wheel.car.index]
        return float(F32(b[1 + self.wheel.k, 2]) - F32(b[0, 2]))


class _Wheel(_Body):
    def __init__(self, car, k):
        _Body.__init__(self, car, 1 + k)
        self.k = k
        self.tiles = set()
        self.color = WHEEL_COLOR
        self.userData = self
        self.joint = _Joint(self)

    def _w(self):
        return self.car.world.ow.wheels()[self.car.index, self.k]

    omega = property(lambda s: float(s._w()[0]))
    phase = property(lambda s: float(s._w()[1]))

    def _fixture(self, f):
        return _FixtureView(self, None)

    @property
    def fixtures(self):
        return [_FixtureView(self, [tuple(v) for v in self.car.world.ow.shape(4)])]


class _Hull(_Body):
    def __init__(self, car):
        _Body.__init__(self, car, 0)
        self.color = (0.8, 0.0, 0.0)
        self.userData = None

    def _fixture(self, f):
        return _FixtureView(self, None)

    @property
    def fixtures(self):   # b2Body.fixtures walks the fixture list newest first
        return [_FixtureView(self, [tuple(v) for v in self.car.world.ow.shape(f)]) for f in (3, 2, 1, 0)]


class Car:
    def __init__(self, world, init_angle, init_x, init_y):
        self.world = world
        self.init_angle, self.init_x, self.init_y = float(init_angle), float(init_x), float(init_y)
        self.index = None
        self.hull = _Hull(self)
        self.wheels = [_Wheel(self, k) for k in range(4)]
        self.drawlist = self.wheels + [self.hull]
        self.fuel_spent = 0.0
        world._register_car(self)

    def _L(self):
        self.world._ensure()
        return self.world.ow.L, self.world.ow.h

    def gas(self, gas):
        L, h = self._L()
        L.orc_ext_gas(h, self.index, float(gas))

    def brake(self, b):
        L, h = self._L()
        L.orc_ext_brake(h, self.index, float(b))

    def steer(self, s):
        L, h = self._L()
        L.orc_ext_steer(h, self.index, float(s))

    def step(self, dt):
        L, h = self._L()
        nt = np.array([len(w.tiles) for w in self.wheels], np.int32)
        L.orc_ext_car_step(h, self.index, nt.ctypes.data, float(dt))

    def draw(self, viewer, draw_particles=True):
        # gym 0.17.2 Car.draw (particles are never drawn in 'state_pixels')
        assert not draw_particles
        for obj in self.drawlist:
            for f in obj.fixtures:
                trans = f.body.transform
                path = [trans * v for v in f.shape.vertices]
                viewer.draw_polygon(path, color=obj.color)
                if "phase" not in dir(obj):
                    continue
                a1 = obj.phase
                a2 = obj.phase + 1.2
                s1, s2, c1, c2 = math.sin(a1), math.sin(a2), math.cos(a1), math.cos(a2)
                if s1 > 0 and s2 > 0:
                    continue
                if s1 > 0:
                    c1 = np.sign(c1)
                if s2 > 0:
                    c2 = np.sign(c2)
                white_poly = [(-WHEEL_W * SIZE, +WHEEL_R * c1 * SIZE), (+WHEEL_W * SIZE, +WHEEL_R * c1 * SIZE),
                              (+WHEEL_W * SIZE, +WHEEL_R * c2 * SIZE), (-WHEEL_W * SIZE, +WHEEL_R * c2 * SIZE)]
                viewer.draw_polygon([trans * v for v in white_poly], color=WHEEL_WHITE)

    def destroy(self):
        self.world._unregister_car(self)


# ------------------------------------------------------------------------------------------
# pyglet.gl + gym rendering
# ------------------------------------------------------------------------------------------
class _Window:
    def __init__(self, w, h):
        self.canvas = np.zeros((96, 96, 3), np.uint8)
        self.context = types.SimpleNamespace()
        self.vp = (96, 96)

    def set_caption(self, s):
        pass

    def switch_to(self):
        _state["window"] = self

    def dispatch_events(self):
        pass

    def clear(self):
        self.canvas[...] = 0

    def flip(self):
        pass

    def close(self):
        pass


class _GL:
    GL_QUADS, GL_POLYGON, GL_TRIANGLES, GL_BLEND = 7, 9, 4, 0x0BE2

    def __init__(self):
        self.mat = None            # (ftx, fty, fdeg, fzoom) while a Transform is enabled
        self.pending = {}
        self.color = (F32(1), F32(1), F32(1))
        self.mode, self.verts = None, []

    def glViewport(self, x, y, w, h):
        assert (w, h) == (96, 96), "only the 96x96 state viewport is stubbed"

    def glPushMatrix(self):
        self.pending = {}

    def glPopMatrix(self):
        self.mat = None

    def glTranslatef(self, x, y, z):
        self.pending["t"] = (F32(x), F32(y))

    def glRotatef(self, deg, x, y, z):
        self.pending["r"] = F32(deg)

    def glScalef(self, sx, sy, sz):
        assert sx == sy
        t, r = self.pending["t"], self.pending["r"]
        fzoom = F32(sx)
        rad = float(r) * (3.14159265358979323846 / 180.0)
        cs, sn = math.cos(rad), math.sin(rad)
        SX, SY = 96.0 / 1000.0, 96.0 / 800.0
        self.mat = (F32(cs * float(fzoom) * SX), F32(-sn * float(fzoom) * SX), F32(float(t[0]) * SX),
                    F32(sn * float(fzoom) * SY), F32(cs * float(fzoom) * SY), F32(float(t[1]) * SY))

    def glColor4f(self, r, g, b, a):
        self.color = (F32(r), F32(g), F32(b))

    def glBegin(self, mode):
        self.mode, self.verts = mode, []

    def glVertex3f(self, x, y, z):
        self.verts.append((F32(x), F32(y), self.color))
        if self.mode == self.GL_QUADS and len(self.verts) == 4:
            self._emit()

    def glEnd(self):
        if self.verts:
            self._emit()
        self.mode = None

    def _project(self, x, y):
        if self.mat is not None:
            m = self.mat
            return F32(F32(m[0] * x) + F32(m[1] * y)) + m[2], F32(F32(m[3] * x) + F32(m[4] * y)) + m[5]
        return x * F32(96.0 / 1000.0), y * F32(96.0 / 800.0)

    def _emit(self, rgb_u8=None):
        n = len(self.verts)
        px = np.array([self._project(v[0], v[1])[0] for v in self.verts], np.float32)
        py = np.array([self._project(v[0], v[1])[1] for v in self.verts], np.float32)
        img = _state["window"].canvas
        L = mo.lib()
        if rgb_u8 is None:
            c = self.verts[0][2]
            L.orc_raster_fill.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_float]
            L.orc_raster_fill(img.ctypes.data, px.ctypes.data, py.ctypes.data, n, float(c[0]), float(c[1]), float(c[2]))
        else:
            L.orc_raster_fill_u8.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int]
            L.orc_raster_fill_u8(img.ctypes.data, px.ctypes.data, py.ctypes.data, n, *rgb_u8)
        self.verts = []


gl = _GL()


class _FilledPolygon:
    def __init__(self, v, color):
        self.v, self.color = v, color

    def render(self):   # gym rendering.Geom.render -> Color.enable -> FilledPolygon.render1
        gl.glColor4f(self.color[0], self.color[1], self.color[2], 1.0)
        if len(self.v) == 4:
            gl.glBegin(gl.GL_QUADS)
        elif len(self.v) > 4:
            gl.glBegin(gl.GL_POLYGON)
        else:
            gl.glBegin(gl.GL_TRIANGLES)
        for p in self.v:
            gl.glVertex3f(p[0], p[1], 0)
        gl.glEnd()


class Viewer:
    def __init__(self, width, height, display=None):
        self.window = _Window(width, height)
        self.isopen = True
        self.onetime_geoms = []

    def draw_polygon(self, v, filled=True, **attrs):
        geom = _FilledPolygon(v, attrs.get("color", (0, 0, 0)))
        self.onetime_geoms.append(geom)
        return geom

    def close(self):
        self.isopen = False


class Transform:
    def __init__(self, translation=(0.0, 0.0), rotation=0.0, scale=(1, 1)):
        self.set_translation(*translation)
        self.set_rotation(rotation)
        self.set_scale(*scale)

    def enable(self):
        gl.glPushMatrix()
        gl.glTranslatef(self.translation[0], self.translation[1], 0)
        gl.glRotatef(57.29577951308232 * self.rotation, 0, 0, 1.0)
        gl.glScalef(self.scale[0], self.scale[1], 1)

    def disable(self):
        gl.glPopMatrix()

    def set_translation(self, newx, newy):
        self.translation = (float(newx), float(newy))

    def set_rotation(self, new):
        self.rotation = float(new)

    def set_scale(self, newx, newy):
        self.scale = (float(newx), float(newy))


class _Label:
    def __init__(self, text="", **kw):
        self.text = text

    def draw(self):
        L = mo.lib()
        L.orc_raster_text.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_raster_text(_state["window"].canvas.ctypes.data, self.text.encode())


def _graphics_draw(count, mode, vdata, cdata):
    assert mode == gl.GL_TRIANGLES and vdata[0] == 'v2i' and cdata[0] == 'c3B'
    v = vdata[1]
    gl.verts = [(F32(v[2 * i]), F32(v[2 * i + 1]), None) for i in range(count)]
    gl._emit(rgb_u8=tuple(int(c) for c in cdata[1][:3]))


class _ImageData:
    def get_data(self, *a):
        c = _state["window"].canvas
        rgba = np.concatenate([c[::-1], np.full((96, 96, 1), 255, np.uint8)], axis=2)   # GL rows: bottom first
        return rgba.tobytes()


class _ColorBuffer:
    def get_image_data(self):
        return _ImageData()


class _BufferManager:
    def get_color_buffer(self):
        return _ColorBuffer()


# ------------------------------------------------------------------------------------------
# gym core / spaces / seeding, shapely
# ------------------------------------------------------------------------------------------
class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.asarray(low).shape


class _Env:
    pass


class _EzPickle:
    def __init__(self, *a, **kw):
        pass


class _Point:
    def __init__(self, xy):
        self.xy = (float(xy[0]), float(xy[1]))

    def within(self, poly):
        """shapely Point.within(Polygon) for the convex road / border quads: strictly inside, float64."""
        px, py = self.xy
        pos = neg = 0
        n = len(poly.pts)
        for k in range(n):
            x0, y0 = poly.pts[k]
            x1, y1 = poly.pts[(k + 1) % n]
            cr = (x1 - x0) * (py - y0) - (y1 - y0) * (px - x0)
            if cr > 0:
                pos += 1
            elif cr < 0:
                neg += 1
        return pos == n or neg == n


class _Polygon:
    def __init__(self, pts):
        self.pts = [(float(p[0]), float(p[1])) for p in pts]


def install():
    """Put the stub modules into sys.modules so `import gym_multi_car_racing` works."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    b2 = mod("Box2D.b2", edgeShape=edgeShape, circleShape=circleShape, fixtureDef=fixtureDef, polygonShape=polygonShape,
             revoluteJointDef=revoluteJointDef, contactListener=contactListener)
    mod("Box2D", b2World=b2World, b2=b2)
    cd = mod("gym.envs.box2d.car_dynamics", Car=Car, SIZE=SIZE, WHEEL_W=WHEEL_W, WHEEL_R=WHEEL_R, WHEELPOS=WHEELPOS)
    box2d = mod("gym.envs.box2d", car_dynamics=cd)
    rendering = mod("gym.envs.classic_control.rendering", Viewer=Viewer, Transform=Transform)
    cc = mod("gym.envs.classic_control", rendering=rendering)
    registration = mod("gym.envs.registration", register=lambda **kw: None)
    envs = mod("gym.envs", box2d=box2d, classic_control=cc, registration=registration)
    spaces = mod("gym.spaces", Box=_Box)
    seeding = mod("gym.utils.seeding", np_random=mo.np_random)
    utils = mod("gym.utils", colorize=lambda s, *a, **k: s, seeding=seeding, EzPickle=_EzPickle)
    mod("gym", Env=_Env, spaces=spaces, utils=utils, envs=envs)
    pgl = mod("pyglet.gl")
    for name in dir(gl):
        if name.startswith("gl") or name.startswith("GL_"):
            setattr(pgl, name, getattr(gl, name))
    text = mod("pyglet.text", Label=_Label)
    graphics = mod("pyglet.graphics", draw=_graphics_draw)
    image = mod("pyglet.image", get_buffer_manager=lambda: _BufferManager())
    window = mod("pyglet.window", key=types.SimpleNamespace())
    mod("pyglet", gl=pgl, text=text, graphics=graphics, image=image, window=window)
    geom = mod("shapely.geometry", Point=_Point, Polygon=_Polygon)
    mod("shapely", geometry=geom)
    # numpy >= 2 removed the binary mode of np.fromstring that the reference uses at :600
    np.fromstring = lambda s, dtype=float, sep='': np.frombuffer(s, dtype=dtype).copy()
