"""Host-side float32 vector / matrix helpers.

Mirror of the reference's `vector.go`, `matrix.go` and `math32.go` for the
values that stay on the host: the per-object world / view /
perspective / screen matrices are built here, with the reference's float32
operation order, and handed to the CUDA path as finished row-major 4x4 f32
arrays.  Everything is `numpy.float32` scalar arithmetic (IEEE-754 binary32,
round-to-nearest-even, no fused multiply-add), so the results are the ones
Go's amd64 build produces, up to the host libm's sin/cos/tan.

Matrices are `numpy.ndarray` of shape (4, 4), dtype float32, row-major — the
memory layout of the reference's `Matrix [4][4]float32` (matrix.go:3).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

f32 = np.float32
pi32 = f32(math.pi)  # math32.go:7-9

_ZERO = f32(0.0)
_ONE = f32(1.0)


def sqrt32(x) -> np.float32:
    """math32.go:11-13."""
    return f32(math.sqrt(float(x)))


def sin32(x) -> np.float32:
    """math32.go:15-17."""
    return f32(math.sin(float(x)))


def cos32(x) -> np.float32:
    """math32.go:19-21."""
    return f32(math.cos(float(x)))


def tan32(x) -> np.float32:
    """math32.go:23-25."""
    return f32(math.tan(float(x)))


# ---------------------------------------------------------------- Vec3 (vector.go:35-85)

def vec3(x, y, z) -> np.ndarray:
    return np.array([x, y, z], dtype=np.float32)


def vec3_sub(a, b) -> np.ndarray:
    return vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2])


def vec3_cross(a, b) -> np.ndarray:
    """vector.go:67-72."""
    x = a[1] * b[2] - a[2] * b[1]
    y = a[2] * b[0] - a[0] * b[2]
    z = a[0] * b[1] - a[1] * b[0]
    return vec3(x, y, z)


def vec3_dot(a, b) -> np.float32:
    """vector.go:74-76 (left-to-right)."""
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def vec3_length(a) -> np.float32:
    """vector.go:63-65."""
    return sqrt32(f32(f32(f32(a[0] * a[0]) + f32(a[1] * a[1])) + f32(a[2] * a[2])))


def vec3_normalize(a) -> np.ndarray:
    """vector.go:78-80: true division by the length."""
    a = np.asarray(a, dtype=np.float32)
    with np.errstate(all="ignore"):
        n = vec3_length(a)
        return vec3(a[0] / n, a[1] / n, a[2] / n)


def vec3_to_radians(a) -> np.ndarray:
    """vector.go:82-85."""
    a = np.asarray(a, dtype=np.float32)
    f = f32(pi32 / f32(180))
    return vec3(a[0] * f, a[1] * f, a[2] * f)


# ---------------------------------------------------------------- Matrix (matrix.go)

def _m(rows: Sequence[Sequence[float]]) -> np.ndarray:
    return np.array(rows, dtype=np.float32)


def NewIdentityMatrix() -> np.ndarray:
    """matrix.go:5-12."""
    return np.eye(4, dtype=np.float32)


def NewScaleMatrix(x, y, z) -> np.ndarray:
    """matrix.go:14-21."""
    return _m([[x, 0, 0, 0], [0, y, 0, 0], [0, 0, z, 0], [0, 0, 0, 1]])


def NewTranslationMatrix(x, y, z) -> np.ndarray:
    """matrix.go:23-30."""
    return _m([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]])


def NewRotationXMatrix(angle) -> np.ndarray:
    """matrix.go:32-44 (angle == 0 short-circuits to identity)."""
    angle = f32(angle)
    if angle == 0:
        return NewIdentityMatrix()
    s, c = sin32(angle), cos32(angle)
    return _m([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]])


def NewRotationYMatrix(angle) -> np.ndarray:
    """matrix.go:46-58."""
    angle = f32(angle)
    if angle == 0:
        return NewIdentityMatrix()
    s, c = sin32(angle), cos32(angle)
    return _m([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]])


def NewRotationZMatrix(angle) -> np.ndarray:
    """matrix.go:60-72."""
    angle = f32(angle)
    if angle == 0:
        return NewIdentityMatrix()
    s, c = sin32(angle), cos32(angle)
    return _m([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def Multiply(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Matrix.Multiply (matrix.go:155-165): res starts at 0, k = 0..3 in order."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    res = np.zeros((4, 4), dtype=np.float32)
    with np.errstate(all="ignore"):
        for k in range(4):
            # res[i][j] += a[i][k] * b[k][j]; float32 array ops round each step
            res = res + a[:, k:k + 1] * b[k:k + 1, :]
    return res


def NewRotationMatrix(x, y, z) -> np.ndarray:
    """matrix.go:74-80."""
    m = NewIdentityMatrix()
    m = Multiply(m, NewRotationXMatrix(x))
    m = Multiply(m, NewRotationYMatrix(y))
    m = Multiply(m, NewRotationZMatrix(z))
    return m


def NewWorldMatrix(scale, rotation, translation) -> np.ndarray:
    """matrix.go:82-88: T * (R * (S * I))."""
    m = NewIdentityMatrix()
    m = Multiply(NewScaleMatrix(scale[0], scale[1], scale[2]), m)
    m = Multiply(NewRotationMatrix(rotation[0], rotation[1], rotation[2]), m)
    m = Multiply(NewTranslationMatrix(translation[0], translation[1], translation[2]), m)
    return m


def NewPerspectiveMatrix(fov, aspect, zNear, zFar) -> np.ndarray:
    """matrix.go:92-106."""
    fov, aspect, zNear, zFar = f32(fov), f32(aspect), f32(zNear), f32(zFar)
    with np.errstate(all="ignore"):
        tanHalfFov = tan32(f32(fov / f32(2.0)))
        m00 = f32(_ONE / f32(aspect * tanHalfFov))
        m11 = f32(_ONE / tanHalfFov)
        m22 = f32(f32(zFar + zNear) / f32(zNear - zFar))
        m23 = f32(f32(f32(f32(2) * zFar) * zNear) / f32(zNear - zFar))
    return _m([[m00, 0, 0, 0], [0, m11, 0, 0], [0, 0, -m22, -m23], [0, 0, -1, 0]])


def NewScreenMatrix(width: int, height: int) -> np.ndarray:
    """matrix.go:108-118 (no Y flip)."""
    hw = f32(f32(width) / f32(2))
    hh = f32(f32(height) / f32(2))
    return _m([[hw, 0, 0, hw], [0, hh, 0, hh], [0, 0, 0.5, 0.5], [0, 0, 0, 1]])


def _view_from_axes(x, y, z, eye) -> np.ndarray:
    return _m([
        [x[0], x[1], x[2], -vec3_dot(x, eye)],
        [y[0], y[1], y[2], -vec3_dot(y, eye)],
        [z[0], z[1], z[2], -vec3_dot(z, eye)],
        [0, 0, 0, 1],
    ])


def NewLookAtMatrix(eye, target, up) -> np.ndarray:
    """matrix.go:120-131."""
    eye = np.asarray(eye, dtype=np.float32)
    z = vec3_normalize(vec3_sub(np.asarray(target, dtype=np.float32), eye))
    x = vec3_normalize(vec3_cross(np.asarray(up, dtype=np.float32), z))
    y = vec3_normalize(vec3_cross(z, x))
    return _view_from_axes(x, y, z, eye)


def NewViewMatrix(eye, direction, up) -> np.ndarray:
    """matrix.go:133-144."""
    eye = np.asarray(eye, dtype=np.float32)
    z = vec3_normalize(np.asarray(direction, dtype=np.float32))
    x = vec3_normalize(vec3_cross(np.asarray(up, dtype=np.float32), z))
    y = vec3_normalize(vec3_cross(z, x))
    return _view_from_axes(x, y, z, eye)


def Transpose(m: np.ndarray) -> np.ndarray:
    """matrix.go:146-153."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T)


def mvp_matrix(perspective, view, world) -> np.ndarray:
    """renderer.go:259-262: ((I * P) * V) * W."""
    m = NewIdentityMatrix()
    m = Multiply(m, perspective)
    m = Multiply(m, view)
    m = Multiply(m, world)
    return m


def light_direction() -> np.ndarray:
    """renderer.go:265: Vec3{-1, 1, 1}.Normalize()."""
    return vec3_normalize(vec3(-1, 1, 1))
