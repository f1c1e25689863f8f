"""Restatement of the handful of ``rowan`` (pin ``>=1.2,<1.3``, reference ``setup.py:41``) functions
that carry arithmetic on the sdanalysis hot path.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

rowan is pure numpy; it is not installed here, so its published algorithm is restated [recalled].
Reference call sites: ``divide``/``to_euler`` ``dynamics/_util.py:151-153``; ``from_euler``
``dynamics/dynamics.py:310-314`` and ``util.py:157``; ``rotate`` ``util.py:143``;
``geometry.intrinsic_distance`` ``order.py:63-65``.

dtype behaviour that matters for parity (all follows from numpy promotion in rowan's own code):

* ``inverse`` copies its input with ``dtype=float`` -> everything downstream of ``divide`` is f64
  even for f32 quaternions.
* ``multiply`` allocates an f64 output but evaluates the products in the dtype of its inputs, so
  ``multiply(f32, f32)`` has f32-rounded products stored in f64 (this is what
  ``intrinsic_distance`` sees in ``order._neighbour_relative_angle``).
* ``from_axis_angle`` halves the angle and takes cos/sin in the angle's dtype; the (integer) basis
  axis is normalised in f64, so the vector part is ``f64(axis) * sin_f32`` and the scalar part
  ``cos_f32`` upcast by the concatenate.
"""

import numpy as np

__all__ = [
    "norm",
    "conjugate",
    "inverse",
    "multiply",
    "divide",
    "from_axis_angle",
    "from_euler",
    "to_matrix",
    "to_euler",
    "rotate",
    "log",
    "intrinsic_distance",
]


def norm(q):
    return np.linalg.norm(np.asarray(q), axis=-1)


def conjugate(q):
    out = np.array(q)
    out[..., 1:] *= -1
    return out


def inverse(q):
    inv = np.array(q, dtype=float)
    normsq = norm(inv) ** 2
    if np.any(normsq):
        inv[..., 1:] *= -1
        nz = normsq > 0
        inv[nz] = inv[nz] / normsq[nz, np.newaxis]
    return inv


def multiply(qi, qj):
    qi = np.asarray(qi)
    qj = np.asarray(qj)
    out = np.empty(np.broadcast(qi, qj).shape)
    out[..., 0] = qi[..., 0] * qj[..., 0] - np.sum(qi[..., 1:] * qj[..., 1:], axis=-1)
    out[..., 1:] = (
        qi[..., 0, np.newaxis] * qj[..., 1:]
        + qj[..., 0, np.newaxis] * qi[..., 1:]
        + np.cross(qi[..., 1:], qj[..., 1:])
    )
    return out


def divide(qi, qj):
    return multiply(qi, inverse(qj))


def from_axis_angle(axes, angles):
    axes = np.asarray(axes)
    angles = np.asarray(angles)
    if axes.shape[:-1] != angles.shape:
        shape = np.broadcast(axes[..., 0], angles).shape
        axes = np.broadcast_to(axes, shape + (3,))
        angles = np.broadcast_to(angles, shape)
    norms = np.linalg.norm(axes, axis=-1)
    axes = axes / norms[..., np.newaxis]
    half = angles / 2.0
    return np.concatenate(
        (np.cos(half)[..., np.newaxis], axes * np.sin(half)[..., np.newaxis]), axis=-1
    )


_BASIS = {"x": np.array([1, 0, 0]), "y": np.array([0, 1, 0]), "z": np.array([0, 0, 1])}


def from_euler(alpha, beta, gamma, convention="zyx", axis_type="intrinsic"):
    """Default convention: q = q_z(alpha) (x) q_y(beta) (x) q_x(gamma)."""
    angles = np.broadcast_arrays(alpha, beta, gamma)
    quats = [from_axis_angle(_BASIS[c], a) for c, a in zip(convention.lower(), angles)]
    if axis_type == "intrinsic":
        return multiply(quats[0], multiply(quats[1], quats[2]))
    return multiply(quats[2], multiply(quats[1], quats[0]))


def _validate_unit(q):
    if not np.allclose(norm(q), 1.0):
        raise ValueError("Arguments must be unit quaternions")


def to_matrix(q, require_unit=True):
    q = np.asarray(q)
    s = norm(q)
    if np.any(s == 0.0):
        raise ZeroDivisionError("At least one element of q has approximately zero norm")
    if require_unit and not np.allclose(s, 1.0):
        raise ValueError("Not all quaternions in q are unit quaternions")
    m = np.empty(q.shape[:-1] + (3, 3))
    # rowan writes ``s **= -1.0``: the scale is 1/|q| (not 1/|q|^2).
    s = 1.0 / s
    w, x, y, z = (q[..., i] for i in range(4))
    m[..., 0, 0] = 1.0 - 2 * s * (y ** 2 + z ** 2)
    m[..., 0, 1] = 2 * s * (x * y - z * w)
    m[..., 0, 2] = 2 * s * (x * z + y * w)
    m[..., 1, 0] = 2 * s * (x * y + z * w)
    m[..., 1, 1] = 1.0 - 2 * s * (x ** 2 + z ** 2)
    m[..., 1, 2] = 2 * s * (y * z - x * w)
    m[..., 2, 0] = 2 * s * (x * z - y * w)
    m[..., 2, 1] = 2 * s * (y * z + x * w)
    m[..., 2, 2] = 1.0 - 2 * s * (x ** 2 + y ** 2)
    return m


def to_euler(q, convention="zyx", axis_type="intrinsic"):
    """Only the default zyx/intrinsic convention (the one the reference uses) is restated.

    Returns (alpha, beta, gamma) with alpha the angle about z (pinned by reference
    ``util.py:150-167`` and ``test/util_test.py:219-250``).
    """
    if convention != "zyx" or axis_type != "intrinsic":
        raise NotImplementedError("oracle restates only the default convention")
    q = np.asarray(q)
    _validate_unit(q)
    atol = 1e-3
    mats = np.clip(to_matrix(q), -1, 1)
    with np.errstate(invalid="ignore"):
        beta = np.arcsin(-mats[..., 2, 0])
    where_zero = np.isclose(np.cos(beta), 0, atol=atol)
    gamma = np.arctan2(mats[..., 2, 1], mats[..., 2, 2])
    alpha = np.arctan2(mats[..., 1, 0], mats[..., 0, 0])
    zero_terms = np.arctan2(-mats[..., 0, 1], mats[..., 1, 1])
    gamma = np.where(where_zero, 0.0, gamma)
    alpha = np.where(where_zero, zero_terms, alpha)
    return np.stack((alpha, beta, gamma), axis=-1)


def _promote_vec(v):
    v = np.asarray(v)
    return np.concatenate((np.zeros(v.shape[:-1] + (1,)), v), axis=-1)


def rotate(q, v):
    q = np.asarray(q)
    return multiply(q, multiply(_promote_vec(v), conjugate(q)))[..., 1:]


def log(q):
    q = np.asarray(q)
    logs = np.empty(q.shape)
    q_norms = norm(q)
    zero = q_norms == 0
    logs[zero, 0] = -np.inf
    logs[zero, 1:] = 0
    nz = ~zero
    v_norms = np.linalg.norm(q[..., 1:], axis=-1)
    v_zero = np.isclose(v_norms, 0)
    sel = nz & v_zero
    logs[sel, 0] = np.log(q_norms[sel])
    logs[sel, 1:] = 0
    sel = nz & ~v_zero
    logs[sel, 0] = np.log(q_norms[sel])
    with np.errstate(invalid="ignore"):
        ang = np.arccos(np.clip(q[sel, 0] / q_norms[sel], -1, 1))
    logs[sel, 1:] = (q[sel, 1:] / v_norms[sel, np.newaxis]) * ang[..., np.newaxis]
    return logs


def intrinsic_distance(p, q):
    """||ln(p^-1 q)|| for unit quaternions, range [0, pi] (reference ``order.py:61-65``)."""
    return norm(log(multiply(conjugate(p), q)))
