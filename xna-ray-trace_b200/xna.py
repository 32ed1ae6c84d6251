"""Host-side XNA 4.0 value-type math in strict binary32 (numpy float32 scalars).

In the reference these values are produced on the C# host by Microsoft.Xna.Framework
(`Camera.cs:40-54`, `SceneObject.cs:183-199`) and would keep being produced there; this module
is the host mirror used by the Python harness so that the same matrices reach libxnart.so and
the oracle.  Evaluation order follows SURVEY.md Appendix A.1 (each operator rounded to float32,
no fused operations — numpy float32 scalar arithmetic rounds after every operator).
"""
import math

import numpy as np

F = np.float32


def v3(x, y, z):
    return np.array([x, y, z], dtype=np.float32)


def dot(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def cross(a, b):
    return np.array(
        [F(F(a[1] * b[2]) - F(a[2] * b[1])), F(F(a[2] * b[0]) - F(a[0] * b[2])), F(F(a[0] * b[1]) - F(a[1] * b[0]))],
        dtype=np.float32,
    )


def normalize(v):
    s = F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))
    with np.errstate(divide="ignore", invalid="ignore"):
        f = F(1.0) / F(math.sqrt(float(s)))
        return np.array([F(v[0] * f), F(v[1] * f), F(v[2] * f)], dtype=np.float32)


def identity():
    return np.eye(4, dtype=np.float32)


def multiply(a, b):
    """Matrix.Multiply: (((a_r1*b_1c) + (a_r2*b_2c)) + (a_r3*b_3c)) + (a_r4*b_4c)."""
    a = np.asarray(a, dtype=np.float32).reshape(4, 4)
    b = np.asarray(b, dtype=np.float32).reshape(4, 4)
    c = np.zeros((4, 4), dtype=np.float32)
    for r in range(4):
        for k in range(4):
            c[r, k] = F(F(F(F(a[r, 0] * b[0, k]) + F(a[r, 1] * b[1, k])) + F(a[r, 2] * b[2, k])) + F(a[r, 3] * b[3, k]))
    return c


def transform(p, m):
    """Vector3.Transform (row vector x matrix)."""
    m = np.asarray(m, dtype=np.float32).reshape(4, 4)
    out = np.zeros(3, dtype=np.float32)
    for k in range(3):
        out[k] = F(F(F(F(p[0] * m[0, k]) + F(p[1] * m[1, k])) + F(p[2] * m[2, k])) + m[3, k])
    return out


def invert(m):
    """Matrix.Invert (cofactor expansion, Appendix A.1)."""
    m = np.asarray(m, dtype=np.float32).reshape(4, 4)
    a11, a12, a13, a14 = m[0]
    a21, a22, a23, a24 = m[1]
    a31, a32, a33, a34 = m[2]
    a41, a42, a43, a44 = m[3]

    def d2(p, q, r, s):  # (p*q) - (r*s)
        return F(F(p * q) - F(r * s))

    def c3(p, x, q, y, r, z):  # ((p*x) - (q*y)) + (r*z)
        return F(F(F(p * x) - F(q * y)) + F(r * z))

    s34_34, s34_24, s34_23 = d2(a33, a44, a34, a43), d2(a32, a44, a34, a42), d2(a32, a43, a33, a42)
    s34_14, s34_13, s34_12 = d2(a31, a44, a34, a41), d2(a31, a43, a33, a41), d2(a31, a42, a32, a41)
    c1 = c3(a22, s34_34, a23, s34_24, a24, s34_23)
    c2 = F(-c3(a21, s34_34, a23, s34_14, a24, s34_13))
    c3_ = c3(a21, s34_24, a22, s34_14, a24, s34_12)
    c4 = F(-c3(a21, s34_23, a22, s34_13, a23, s34_12))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = F(1.0) / F(F(F(F(a11 * c1) + F(a12 * c2)) + F(a13 * c3_)) + F(a14 * c4))
        r = np.zeros((4, 4), dtype=np.float32)
        r[0, 0], r[1, 0], r[2, 0], r[3, 0] = F(c1 * inv), F(c2 * inv), F(c3_ * inv), F(c4 * inv)
        r[0, 1] = F(F(-c3(a12, s34_34, a13, s34_24, a14, s34_23)) * inv)
        r[1, 1] = F(c3(a11, s34_34, a13, s34_14, a14, s34_13) * inv)
        r[2, 1] = F(F(-c3(a11, s34_24, a12, s34_14, a14, s34_12)) * inv)
        r[3, 1] = F(c3(a11, s34_23, a12, s34_13, a13, s34_12) * inv)
        s24_34, s24_24, s24_23 = d2(a23, a44, a24, a43), d2(a22, a44, a24, a42), d2(a22, a43, a23, a42)
        s24_14, s24_13, s24_12 = d2(a21, a44, a24, a41), d2(a21, a43, a23, a41), d2(a21, a42, a22, a41)
        r[0, 2] = F(c3(a12, s24_34, a13, s24_24, a14, s24_23) * inv)
        r[1, 2] = F(F(-c3(a11, s24_34, a13, s24_14, a14, s24_13)) * inv)
        r[2, 2] = F(c3(a11, s24_24, a12, s24_14, a14, s24_12) * inv)
        r[3, 2] = F(F(-c3(a11, s24_23, a12, s24_13, a13, s24_12)) * inv)
        s23_34, s23_24, s23_23 = d2(a23, a34, a24, a33), d2(a22, a34, a24, a32), d2(a22, a33, a23, a32)
        s23_14, s23_13, s23_12 = d2(a21, a34, a24, a31), d2(a21, a33, a23, a31), d2(a21, a32, a22, a31)
        r[0, 3] = F(F(-c3(a12, s23_34, a13, s23_24, a14, s23_23)) * inv)
        r[1, 3] = F(c3(a11, s23_34, a13, s23_14, a14, s23_13) * inv)
        r[2, 3] = F(F(-c3(a11, s23_24, a12, s23_14, a14, s23_12)) * inv)
        r[3, 3] = F(c3(a11, s23_23, a12, s23_13, a13, s23_12) * inv)
    return r


def create_look_at(eye, target, up):
    eye, target, up = (np.asarray(a, dtype=np.float32) for a in (eye, target, up))
    z = normalize(eye - target)
    x = normalize(cross(up, z))
    y = cross(z, x)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, :3] = (x[0], y[0], z[0])
    m[1, :3] = (x[1], y[1], z[1])
    m[2, :3] = (x[2], y[2], z[2])
    m[3, :3] = (F(-dot(x, eye)), F(-dot(y, eye)), F(-dot(z, eye)))
    m[3, 3] = 1.0
    return m


def create_perspective_fov(fov, aspect, near, far):
    fov, aspect, near, far = F(fov), F(aspect), F(near), F(far)
    m = np.zeros((4, 4), dtype=np.float32)
    y_scale = F(1.0) / F(math.tan(float(F(fov * F(0.5)))))
    m[0, 0] = F(y_scale / aspect)
    m[1, 1] = y_scale
    m[2, 2] = F(far / F(near - far))
    m[2, 3] = -1.0
    m[3, 2] = F(F(near * far) / F(near - far))
    return m


def create_scale(s):
    m = identity()
    m[0, 0], m[1, 1], m[2, 2] = F(s[0]), F(s[1]), F(s[2])
    return m


def create_translation(p):
    m = identity()
    m[3, 0], m[3, 1], m[3, 2] = F(p[0]), F(p[1]), F(p[2])
    return m


def _cs(r):
    return F(math.cos(float(F(r)))), F(math.sin(float(F(r))))


def create_rotation_x(r):
    c, s = _cs(r)
    m = identity()
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, s, F(-s), c
    return m


def create_rotation_y(r):
    c, s = _cs(r)
    m = identity()
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, F(-s), s, c
    return m


def create_rotation_z(r):
    c, s = _cs(r)
    m = identity()
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, s, F(-s), c
    return m


def aspect_ratio(width, height):
    """Viewport.AspectRatio = (float)Width / (float)Height."""
    return F(F(width) / F(height))


# ---- Color, vectorised over arrays of packed R | G<<8 | B<<16 | A<<24 colours ----
def color_to_vector3(packed):
    """Color.ToVector3(): (R/255f, G/255f, B/255f), each a binary32 division.  (..,) uint32 -> (.., 3) float32."""
    p = np.asarray(packed, dtype=np.uint32)
    ch = np.stack([p & 255, (p >> 8) & 255, (p >> 16) & 255], axis=-1).astype(np.float32)
    return ch / np.float32(255.0)


def color_from_vector3(v, truncate=False):
    """new Color(Vector3), XNA 4.0: clamp(v * 255f) then round half to even (MonoGame: truncate); A = 255."""
    t = np.asarray(v, dtype=np.float32) * np.float32(255.0)
    t = np.where(np.isnan(t), np.float32(0.0), t)
    t = np.clip(t, np.float32(0.0), np.float32(255.0))
    q = (np.floor(t.astype(np.float64)) if truncate else np.rint(t.astype(np.float64))).astype(np.uint32)
    return (q[..., 0] | (q[..., 1] << 8) | (q[..., 2] << 16) | np.uint32(0xFF000000)).astype(np.uint32)


def length3(v):
    """Vector3.Length(): (float)Math.Sqrt(((x*x) + (y*y)) + (z*z)), vectorised over (.., 3) float32."""
    v = np.asarray(v, dtype=np.float32)
    s = ((v[..., 0] * v[..., 0]) + (v[..., 1] * v[..., 1])) + (v[..., 2] * v[..., 2])
    return np.sqrt(s.astype(np.float64)).astype(np.float32)
