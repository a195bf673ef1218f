"""np_oracle.py — second, independent numpy restatement of linon's hot path + scene inputs.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference has no tests or golden vectors and
its Rust host cannot be executed in this image; this file and oracle/linon_oracle.cpp are two
separately written restatements of the same WGSL, cross-checked against each other in
tests/test_oracle_selfcheck.py.  What pins them to the reference is oracle/_ref: the reference's own
shader text transpiled to C++ (oracle/wgsl_ref), bit-identical with linon_oracle.cpp
(tests/test_wgsl_reference_cpu.py).

What lives here
  * scene inputs the reference builds on the host (the product has its own C++ copy that the tests
    compare with this one):
      cornell_box()          src/cornell_box.rs:2-163 (18 quads -> 72 vertices, 36 faces)
      default_scene()        src/application.rs:105-123 (negate x,z; normalise; centre)
      normalize_vertices()   src/vertices.rs:3-17
      get_center()           src/vertices.rs:19-57
      Arcball                src/arcball.rs:33-129 on cgmath 0.18 semantics (Cargo.lock:486-489)
      camera_uniform()       src/main_view.rs:21-31
      default_settings()     src/main_view.rs:133-139
  * a vectorised fp32 march (all rays advance together under masks):
      march()                src/main_view.wgsl:229-279, 339-375
      FIELD_FUNCTIONS        src/functions.rs:36-126
      overlay()              src/overlay.wgsl:99-146
Everything is float32; Python scalars are weak under NEP 50 so `arr32 * 0.5` stays float32.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# ----------------------------------------------------------------------------------------------
# scene
# ----------------------------------------------------------------------------------------------
# The Cornell box of src/cornell_box.rs is 18 quads; each quad (a,b,c,d) contributes the faces
# (a,b,c) and (c,d,a) — exactly the pattern of INDICES (src/cornell_box.rs:102-163).
_CORNELL_QUADS = (
    # floor, short-block footprint, tall-block footprint            (src/cornell_box.rs:4-17)
    ((552.8, 0, 0), (0, 0, 0), (0, 0, 559.2), (549.6, 0, 559.2)),
    ((130, 0, 65), (82, 0, 225), (240, 0, 272), (290, 0, 114)),
    ((423, 0, 247), (265, 0, 296), (314, 0, 456), (472, 0, 406)),
    # ceiling, light                                                 (:20-28)
    ((556, 548.8, 0), (556, 548.8, 559.2), (0, 548.8, 559.2), (0, 548.8, 0)),
    ((343, 548.8, 227), (343, 548.8, 332), (213, 548.8, 332), (213, 548.8, 227)),
    # back, right, left wall                                         (:31-46)
    ((549.6, 0, 559.2), (0, 0, 559.2), (0, 548.8, 559.2), (556, 548.8, 559.2)),
    ((0, 0, 559.2), (0, 0, 0), (0, 548.8, 0), (0, 548.8, 559.2)),
    ((552.8, 0, 0), (549.6, 0, 559.2), (556, 548.8, 559.2), (556, 548.8, 0)),
    # short block                                                    (:49-72)
    ((130, 165, 65), (82, 165, 225), (240, 165, 272), (290, 165, 114)),
    ((290, 0, 114), (290, 165, 114), (240, 165, 272), (240, 0, 272)),
    ((130, 0, 65), (130, 165, 65), (290, 165, 114), (290, 0, 114)),
    ((82, 0, 225), (82, 165, 225), (130, 165, 65), (130, 0, 65)),
    ((240, 0, 272), (240, 165, 272), (82, 165, 225), (82, 0, 225)),
    # tall block                                                     (:75-98)
    ((423, 330, 247), (265, 330, 296), (314, 330, 456), (472, 330, 406)),
    ((423, 0, 247), (423, 330, 247), (472, 330, 406), (472, 0, 406)),
    ((472, 0, 406), (472, 330, 406), (314, 330, 456), (314, 0, 456)),
    ((314, 0, 456), (314, 330, 456), (265, 330, 296), (265, 0, 296)),
    ((265, 0, 296), (265, 330, 296), (423, 330, 247), (423, 0, 247)),
)


def cornell_box():
    """Raw (un-normalised) vertices [72,3] f32 and faces [36,3] u32."""
    verts = np.array(_CORNELL_QUADS, dtype=f32).reshape(-1, 3)
    faces = []
    for q in range(len(_CORNELL_QUADS)):
        a = 4 * q
        faces += [(a, a + 1, a + 2), (a + 2, a + 3, a)]
    return verts, np.array(faces, dtype=np.uint32)


def normalize_vertices(verts):
    """src/vertices.rs:3-17: one global min/max over every coordinate."""
    flat = verts.astype(f32).reshape(-1)
    mx, mn = flat.max(), flat.min()
    out = (flat - mn) / (mx - mn) * f32(2.0) - f32(1.0)
    return out.reshape(-1, 3).astype(f32)


def get_center(verts):
    """src/vertices.rs:19-57: bounding-box centre."""
    lo = verts.min(axis=0).astype(f32)
    hi = verts.max(axis=0).astype(f32)
    return ((lo + hi) / f32(2.0)).astype(f32)


def default_scene():
    """src/application.rs:105-123: negate x and z, normalise, centre."""
    verts, faces = cornell_box()
    verts = verts.copy()
    verts[:, 0] = -verts[:, 0]
    verts[:, 2] = -verts[:, 2]
    verts = normalize_vertices(verts)
    return verts, faces, get_center(verts)


def default_settings():
    """src/main_view.rs:133-139."""
    return dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=0, central_difference_delta=1,
                lyapunov_scaling=50.0)


# ---- cgmath-style fp32 4x4 helpers (column-major: m[c][r]) -------------------------------------
def _mat_identity():
    return np.eye(4, dtype=f32)


def _mat_translation(v):
    m = _mat_identity()
    m[3, 0:3] = np.asarray(v, dtype=f32)  # column 3
    return m


def _mat_mul(a, b):
    """(a*b) in column-major storage m[c][r]:  out[c][r] = sum_k a[k][r] * b[c][k]."""
    out = np.zeros((4, 4), dtype=f32)
    for c in range(4):
        for r in range(4):
            s = f32(0.0)
            for k in range(4):
                s = f32(s + f32(a[k, r] * b[c, k]))
            out[c, r] = s
    return out


def _mat_vec(m, v):
    out = np.zeros(4, dtype=f32)
    for r in range(4):
        s = f32(0.0)
        for k in range(4):
            s = f32(s + f32(m[k, r] * v[k]))
        out[r] = s
    return out


def _det3(m):
    return f32(m[0, 0] * (m[1, 1] * m[2, 2] - m[2, 1] * m[1, 2])
               - m[1, 0] * (m[0, 1] * m[2, 2] - m[2, 1] * m[0, 2])
               + m[2, 0] * (m[0, 1] * m[1, 2] - m[1, 1] * m[0, 2]))


def _mat_invert(m):
    """Adjugate / determinant in fp32 (the structure of cgmath's Matrix4::invert)."""
    m = m.astype(f32)
    cof = np.zeros((4, 4), dtype=f32)
    for c in range(4):
        for r in range(4):
            minor = np.delete(np.delete(m, c, axis=0), r, axis=1)
            sign = f32(-1.0) if (c + r) & 1 else f32(1.0)
            cof[c, r] = f32(sign * _det3(minor))
    det = f32(0.0)
    for c in range(4):
        det = f32(det + f32(m[c, 0] * cof[c, 0]))
    inv_det = f32(1.0) / det
    return (cof.T * inv_det).astype(f32)


def _quat_mul(a, b):
    """Hamilton product, quaternions as (s, x, y, z)."""
    s1, x1, y1, z1 = [f32(t) for t in a]
    s2, x2, y2, z2 = [f32(t) for t in b]
    return np.array([
        s1 * s2 - x1 * x2 - y1 * y2 - z1 * z2,
        s1 * x2 + x1 * s2 + y1 * z2 - z1 * y2,
        s1 * y2 + y1 * s2 + z1 * x2 - x1 * z2,
        s1 * z2 + z1 * s2 + x1 * y2 - y1 * x2,
    ], dtype=f32)


def _quat_to_mat(q):
    s, x, y, z = [f32(t) for t in q]
    x2, y2, z2 = x + x, y + y, z + z
    xx2, xy2, xz2 = x2 * x, x2 * y, x2 * z
    yy2, yz2, zz2 = y2 * y, y2 * z, z2 * z
    sy2, sz2, sx2 = y2 * s, z2 * s, x2 * s
    one = f32(1.0)
    m = np.zeros((4, 4), dtype=f32)
    m[0] = [one - yy2 - zz2, xy2 + sz2, xz2 - sy2, 0]
    m[1] = [xy2 - sz2, one - xx2 - zz2, yz2 + sx2, 0]
    m[2] = [xz2 + sy2, yz2 - sx2, one - xx2 - yy2, 0]
    m[3] = [0, 0, 0, 1]
    return m


class Arcball:
    """src/arcball.rs:18-129 (Shoemake arcball as used by MainView)."""

    def __init__(self, center, zoom_speed, screen):
        self.center = np.asarray(center, dtype=f32)
        self.translation = _mat_translation([0.0, 0.0, -1.0])
        self.center_translation = _mat_invert(_mat_translation(self.center))
        self.rotation = np.array([1, 0, 0, 0], dtype=f32)
        self.zoom_speed = f32(zoom_speed)
        self.inv_screen = [f32(1.0) / f32(screen[0]), f32(1.0) / f32(screen[1])]
        self._update()

    def _update(self):
        self.camera = _mat_mul(_mat_mul(self.translation, _quat_to_mat(self.rotation)), self.center_translation)
        self.inv_camera = _mat_invert(self.camera)

    def zoom(self, amount, elapsed):
        motion = np.array([0.0, 0.0, amount], dtype=f32) * self.zoom_speed * f32(elapsed)
        self.translation = _mat_mul(_mat_translation(motion), self.translation)
        self._update()

    @staticmethod
    def _screen_to_arcball(p):
        dist = f32(p[0] * p[0] + p[1] * p[1])
        if dist <= f32(1.0):
            return np.array([0.0, p[0], p[1], np.sqrt(f32(1.0) - dist)], dtype=f32)
        ln = np.sqrt(dist)
        return np.array([0.0, p[0] / ln, p[1] / ln, 0.0], dtype=f32)

    def rotate(self, mouse_prev, mouse_cur):
        one, two = f32(1.0), f32(2.0)
        clamp = lambda x: min(max(f32(x), -one), one)
        m_cur = (clamp(f32(mouse_cur[0]) * two * self.inv_screen[0] - one),
                 clamp(one - two * f32(mouse_cur[1]) * self.inv_screen[1]))
        m_prev = (clamp(f32(mouse_prev[0]) * two * self.inv_screen[0] - one),
                  clamp(one - two * f32(mouse_prev[1]) * self.inv_screen[1]))
        cur_ball = self._screen_to_arcball(m_cur)
        prev_ball = self._screen_to_arcball(m_prev)
        self.rotation = _quat_mul(_quat_mul(cur_ball, prev_ball), self.rotation)
        self._update()

    def eye_pos(self):
        return self.inv_camera[3, 0:3].copy()

    def _dir(self, v4):
        d = _mat_vec(self.inv_camera, np.array(v4, dtype=f32))[0:3]
        return (d / np.sqrt(f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))).astype(f32)

    def eye_dir(self):
        return self._dir([0, 0, -1, 0])

    def up_dir(self):
        return self._dir([0, 1, 0, 0])


def camera_uniform(cam: Arcball):
    """src/main_view.rs:21-31 -> dict(origin, view_direction, up) of float32[3]."""
    return dict(origin=cam.eye_pos(), view_direction=cam.eye_dir(), up=cam.up_dir())


def default_camera(center, width, height):
    """src/main_view.rs:153-155."""
    cam = Arcball(center, 1.0, [width, height])
    cam.zoom(-1.0, 1.0)
    return cam


def orbit_camera(center, width, height, frame, n_frames=64):
    """The C5 orbit of SURVEY 8(d): frame k = k successive horizontal drags of 360/n_frames degrees."""
    cam = default_camera(center, width, height)
    w2, h2 = f32(width) / f32(2.0), f32(height) / f32(2.0)
    cur_x = f32(w2 + f32(np.sin(np.pi / n_frames)) * w2)
    for _ in range(frame):
        cam.rotate((w2, h2), (cur_x, h2))
    return cam


# ----------------------------------------------------------------------------------------------
# vector helpers on [N,3] float32 arrays
# ----------------------------------------------------------------------------------------------
def _dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def _length(a):
    return np.sqrt(_dot(a, a))


def _normalize(a):
    return a / _length(a)[..., None]


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _wgsl_clamp(e, lo, hi):
    return np.minimum(np.maximum(e, lo), hi)


def _v3(x, y, z, n):
    return np.broadcast_to(np.array([x, y, z], dtype=f32), (n, 3))


# ---- src/main_view.wgsl:78-146 ------------------------------------------------------------------
def rotateZ(v, phi):
    c, s = np.cos(phi), np.sin(phi)
    zero = np.zeros_like(c)
    c0 = np.stack([c, s, zero], -1)
    c1 = np.stack([-s, c, zero], -1)
    c2 = np.stack([zero, zero, zero], -1)
    return c0 * v[..., 0:1] + c1 * v[..., 1:2] + c2 * v[..., 2:3]


def refraction_index(t):
    air_pressure = f32(101325.0)
    c1 = f32(0.0000104)
    c2 = f32(0.00366)
    p10 = f32(1e-10)
    return c1 * air_pressure * (f32(1.0) + air_pressure * (f32(60.1) - f32(0.972) * t) * p10) / (f32(1.0) + c2 * t)


def refraction(t_in, t_out, v_in, n):
    eta_in = refraction_index(t_in)
    eta_out = refraction_index(t_out)
    cosi = _wgsl_clamp(f32(-1.0), f32(1.0), _dot(v_in, n))
    neg = cosi < 0
    cosi = np.where(neg, -cosi, cosi)
    n_ref = np.where(neg[..., None], n, -n)
    eta = eta_in / eta_out
    k = f32(1.0) - eta * eta * (f32(1.0) - cosi * cosi)
    refl = v_in - (f32(2.0) * _dot(n_ref, v_in))[..., None] * n_ref
    with np.errstate(invalid="ignore"):
        refr = eta[..., None] * v_in + (eta * cosi - np.sqrt(k))[..., None] * n_ref
    return np.where((k < 0)[..., None], refl, refr).astype(f32)


def point_plane_distance(p, n, p0):
    d = _dot(p0, n)
    return np.abs(_dot(p, n) - d)


def sigmoid(x):
    return f32(1.0) / (f32(1.0) + np.exp(-x))


# ---- src/functions.rs:36-126 ----------------------------------------------------------------------
def _mirage_spherical(p_prev, p, v0, v, t, sig):
    n = p.shape[0]
    t_env, t_src, max_dist = f32(15.0), f32(200.0), f32(0.25)
    center = _v3(-0.5, 0.5, -0.5, n)
    center_dest = p - center
    normal = _normalize(center_dest)
    dist_in = _length(p_prev - center)
    dist_out = _length(center_dest)
    if sig:
        a_in = sigmoid(dist_in / max_dist * f32(12.0) - f32(6.0))
        a_out = sigmoid(dist_out / max_dist * f32(12.0) - f32(6.0))
    else:
        a_in, a_out = dist_in / max_dist, dist_out / max_dist
    part_in = _wgsl_clamp(f32(0.0), f32(1.0), a_in)
    part_out = _wgsl_clamp(f32(0.0), f32(1.0), a_out)
    t_in = part_in * t_env + (f32(1.0) - part_in) * t_src
    t_out = part_out * t_env + (f32(1.0) - part_out) * t_src
    return refraction(t_in, t_out, v, normal)


def _mirage_plane(p_prev, p, v0, v, t, sig):
    n = p.shape[0]
    t_env, t_src, max_dist = f32(15.0), f32(30.0), f32(0.01)
    plane_p0 = _v3(0.0, 0.1, 0.0, n)
    plane_n = _v3(0.0, 1.0, 0.0, n)
    dist_in = point_plane_distance(p_prev, plane_n, plane_p0)
    dist_out = point_plane_distance(p, plane_n, plane_p0)
    if sig:
        a_in = sigmoid(dist_in / max_dist * f32(12.0) - f32(6.0))
        a_out = sigmoid(dist_out / max_dist * f32(12.0) - f32(6.0))
    else:
        a_in, a_out = dist_in / max_dist, dist_out / max_dist
    part_in = _wgsl_clamp(f32(0.0), f32(1.0), a_in)
    part_out = _wgsl_clamp(f32(0.0), f32(1.0), a_out)
    t_in = part_in * t_env + (f32(1.0) - part_in) * t_src
    t_out = part_out * t_env + (f32(1.0) - part_out) * t_src
    return refraction(t_in, t_out, v, plane_n)


def _translation_x(p_prev, p, v0, v, t):
    return np.stack([v0[:, 0] + f32(0.5) * t, v0[:, 1] + f32(0.0), v0[:, 2] + f32(0.0)], -1)


def _translation_z(p_prev, p, v0, v, t):
    return np.stack([v0[:, 0] + f32(0.0), v0[:, 1] + f32(0.0), v0[:, 2] + f32(0.5) * t], -1)


def _rotation(p_prev, p, v0, v, t):
    return _normalize(v0 + rotateZ(v0, f32(3.141592653589793) * t))


def _lorenz(p_prev, p, v0, v, t):
    rho, sigma, beta = f32(28.0), f32(10.0), f32(8.0 / 3.0)
    return np.stack([sigma * (p[:, 1] - p[:, 0]), p[:, 0] * (rho - p[:, 2]) - p[:, 1],
                     p[:, 0] * p[:, 1] - beta * p[:, 2]], -1)


def _roessler(p_prev, p, v0, v, t):
    a, b, c = f32(0.1), f32(0.1), f32(14.0)
    return np.stack([-p[:, 1] - p[:, 2], p[:, 0] + a * p[:, 1], b + p[:, 2] * (p[:, 0] - c)], -1)


FIELD_FUNCTIONS = {  # ids follow the enum order of src/functions.rs:2-13
    1: lambda *a: _mirage_spherical(*a, sig=False),
    2: lambda *a: _mirage_plane(*a, sig=False),
    3: lambda *a: _mirage_spherical(*a, sig=True),
    4: lambda *a: _mirage_plane(*a, sig=True),
    5: _translation_x,
    6: _translation_z,
    7: _rotation,
    8: _lorenz,
    9: _roessler,
}


# ---- src/main_view.wgsl:150-220 over all faces, vectorised over rays --------------------------------
def ray_color(verts, faces, origin, direction, max_dist):
    """Returns colour [N,4] (alpha = t of the hit, 0 for none) and hit id [N]."""
    n = origin.shape[0]
    eps = f32(0.0000001)
    t_best = np.full(n, -1.0, dtype=f32)
    d1 = np.zeros((n, 3), dtype=f32)
    d2 = np.zeros((n, 3), dtype=f32)
    hit = np.full(n, -1, dtype=np.int32)
    with np.errstate(all="ignore"):
        for i, (fa, fb, fc) in enumerate(faces):
            a, b, c = verts[fa], verts[fb], verts[fc]
            edge1 = np.broadcast_to(b - a, (n, 3))
            edge2 = np.broadcast_to(c - a, (n, 3))
            h = _cross(direction, edge2)
            det = _dot(edge1, h)
            ok = ~((det > -eps) & (det < eps))
            f = f32(1.0) / det
            s = origin - a
            u = f * _dot(s, h)
            ok &= ~((u < 0) | (u > 1))
            q = _cross(s, edge1)
            v = f * _dot(direction, q)
            ok &= ~((v < 0) | (u + v > 1))
            t = f * _dot(edge2, q)
            ok &= t > eps
            t_new = np.where(ok, t, f32(-1.0))
            upd = (t_new > 0) & (t_new < max_dist) & ((t_best < 0) | (t_new < t_best))
            t_best = np.where(upd, t_new, t_best)
            d1 = np.where(upd[:, None], edge1, d1)
            d2 = np.where(upd[:, None], edge2, d2)
            hit = np.where(upd, np.int32(i), hit)
        found = (t_best > 0) & (t_best < max_dist)
        normal = _normalize(_cross(d1, d2))
        intensity = np.maximum(_dot(normal, -direction), f32(0.0))
        intensity = np.where(intensity == 0, np.maximum(_dot(-normal, -direction), f32(0.0)), intensity)
        spec = np.power(intensity, f32(64.0))
        rgb = ((f32(0.01) + intensity + spec)[:, None]) * np.abs(normal)
    color = np.zeros((n, 4), dtype=f32)
    color[found, 0:3] = rgb[found]
    color[found, 3] = t_best[found]
    hit = np.where(found, hit, -1).astype(np.int32)
    return color, hit


def pixel_dirs(camera, W, H, xs, ys_tex):
    """src/main_view.wgsl:341-366 for texture-space pixels (xs, ys_tex)."""
    origin = np.asarray(camera["origin"], dtype=f32)
    view = np.asarray(camera["view_direction"], dtype=f32)
    up = np.asarray(camera["up"], dtype=f32)
    width, height = f32(W), f32(H)
    aspect = width / height
    focal, vh = f32(0.035), f32(0.025)
    vw = aspect * vh
    w = -view
    horizontal = _cross(up, w)
    vertical = _cross(w, horizontal)
    nh, nv = _normalize(horizontal), _normalize(vertical)
    gid_x = xs.astype(f32)
    gid_y = (H - 1 - ys_tex).astype(f32)
    u = gid_x / (width - f32(1.0)) * vw - f32(0.5) * vw
    v = gid_y / (height - f32(1.0)) * vh - f32(0.5) * vh
    s = u[:, None] * nh + v[:, None] * nv + focal * view
    return origin, _normalize(s).astype(f32)


def to_unorm8(x):
    x = np.where(np.isnan(x), 0.0, x)
    return np.rint(np.clip(x, 0.0, 1.0).astype(f32) * f32(255.0)).astype(np.uint8)


def march(verts, faces, camera, settings, fn_id, W, H, xs, ys_tex, h_initial=0.1, adaptive=True):
    """src/main_view.wgsl:229-279 for the listed texture-space pixels.

    Returns dict(color [N,4] f32, rgba8 [N,4], endpoint [N,3], hit_id, accepted, rejected, prehit)."""
    fn = FIELD_FUNCTIONS[fn_id]
    xs = np.asarray(xs)
    ys_tex = np.asarray(ys_tex)
    n = xs.shape[0]
    origin, dirs = pixel_dirs(camera, W, H, xs, ys_tex)
    fw = f32(settings["field_weight"])
    h_initial = f32(h_initial)

    cur_point = np.broadcast_to(origin, (n, 3)).astype(f32).copy()
    cur_dir = dirs.copy()
    start_dir = dirs.copy()
    t = np.zeros(n, dtype=f32)
    last_v = np.zeros((n, 3), dtype=f32)
    last_diff = np.full(n, -1.0, dtype=f32)
    h = np.full(n, h_initial, dtype=f32)
    has_color = np.zeros(n, dtype=bool)
    color = np.zeros((n, 4), dtype=f32)
    hit_id = np.full(n, -1, dtype=np.int32)
    accepted = np.zeros(n, dtype=np.int32)
    rejected = np.zeros(n, dtype=np.int32)
    prehit = np.zeros(n, dtype=np.int32)

    while True:
        act = np.nonzero(t <= f32(5.0))[0]
        if act.size == 0:
            break
        cp, cd, sd, ta, ha = cur_point[act], cur_dir[act], start_dir[act], t[act], h[act]
        k1 = fn(cp, cp, sd, cd, ta)
        k2 = fn(cp, cp + (f32(0.5) * ha)[:, None] * k1, sd, k1, ta + f32(0.5) * ha)
        k3 = fn(cp, cp + (f32(0.5) * ha)[:, None] * k2, sd, k2, ta + f32(0.5) * ha)
        k4 = fn(cp, cp + ha[:, None] * k3, sd, k3, ta + ha)
        v = ((k1 + f32(2.0) * k2 + f32(2.0) * k3 + k4) / f32(6.0)).astype(f32)
        diff = _length(v - last_v[act])
        ld = last_diff[act]
        if adaptive:
            rej = (ld >= 0) & (diff > f32(10.0) * ld) & (ha > f32(0.002))
        else:
            rej = np.zeros(act.size, dtype=bool)
        # rejected lanes: h = 0.001, nothing else changes
        ridx = act[rej]
        h[ridx] = f32(0.001)
        rejected[ridx] += 1
        ok = ~rej
        a = act[ok]
        if a.size == 0:
            continue
        cd, v, diff, ld, ha, cp = cd[ok], v[ok], diff[ok], ld[ok], ha[ok], cp[ok]
        cd = ((f32(1.0) - fw) * cd + fw * v).astype(f32)
        step_dir = cd * ha[:, None]
        need = ~has_color[a]
        if need.any():
            ni = a[need]
            sdn = step_dir[need]
            col, hid = ray_color(verts, faces, cp[need], _normalize(sdn), _length(sdn))
            color[ni] = col
            got = col[:, 3] > 0
            has_color[ni] = got
            hit_id[ni[got]] = hid[got]
            prehit[ni] += 1
        cur_dir[a] = cd
        cur_point[a] = cp + step_dir
        t[a] = ta[ok] + ha
        last_v[a] = v
        restore = (f32(2.0) * diff < ld) & (ha < h_initial)
        h[a[restore]] = h_initial
        last_diff[a] = diff
        accepted[a] += 1

    color[:, 3] = 1.0
    color[~has_color] = np.array([0, 0, 0, 1], dtype=f32)
    return dict(color=color, rgba8=to_unorm8(color), endpoint=cur_point, hit_id=hit_id, accepted=accepted,
                rejected=rejected, prehit=prehit)


# ---- src/overlay.wgsl:99-146 -------------------------------------------------------------------------
def overlay(raycast_rgba8, endpoint, settings, exponents=None):
    """raycast_rgba8 [H,W,4] u8, endpoint [H,W,4|3] f32 -> (final_rgba8, exponents [H,W] f32, stride W)."""
    H, W = raycast_rgba8.shape[:2]
    out = raycast_rgba8.copy()
    exps = np.zeros((H, W), dtype=f32) if exponents is None else exponents.copy()
    if settings["overlay_mode"] != 1:
        return out, exps
    d = int(settings["central_difference_delta"])
    pad = 2 * d
    if W - 2 * pad <= 0 or H - 2 * pad <= 0:
        return out, exps
    E = endpoint[..., 0:3].astype(f32)
    ys, xs = slice(pad, H - pad), slice(pad, W - pad)
    gx = f32(0.5) * (E[ys, pad + d:W - pad + d] - E[ys, pad - d:W - pad - d])
    gy = f32(0.5) * (E[pad + d:H - pad + d, xs] - E[pad - d:H - pad - d, xs])
    m00, m01, m11 = _dot(gx, gx), _dot(gx, gy), _dot(gy, gy)
    a0 = m00 * m11 - m01 * m01  # det
    a1 = -(m00 + m11)
    discrim = a1 * a1 - f32(4.0) * a0
    with np.errstate(all="ignore"):
        root = np.sqrt(np.abs(discrim))
        e0 = np.where(discrim >= 0, (-a1 - root) / f32(2.0), -a1 / f32(2.0))
        e1 = np.where(discrim >= 0, (-a1 + root) / f32(2.0), root / f32(2.0))
        lam = np.sqrt(np.maximum(e0, e1))
        s = np.exp(f32(settings["lyapunov_scaling"]) * lam - f32(5.0)).astype(f32)
    exps[ys, xs] = s
    alpha = (np.minimum(f32(1.0), s) * f32(0.5))[..., None]
    col = raycast_rgba8[ys, xs].astype(f32) / f32(255.0)
    blended = to_unorm8(alpha * f32(1.0) + (f32(1.0) - alpha) * col)
    out[ys, xs] = np.where((s >= f32(0.8))[..., None], blended, raycast_rgba8[ys, xs])
    return out, exps
