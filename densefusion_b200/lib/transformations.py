"""The two pose helpers of the reference's vendored transformations.py that the path uses
(lib/transformations.py:1254-1278 and :1281-1363, isprecise branch), float64 numpy on the host.
The on-device equivalents live in csrc/pose.cu; these are for callers that still want numpy."""
from __future__ import annotations

import math

import numpy

_EPS = numpy.finfo(float).eps * 4.0


def quaternion_matrix(quaternion):
    """Homogeneous 4x4 rotation matrix of quaternion (w, x, y, z) of any non-tiny norm."""
    q = numpy.array(quaternion, dtype=numpy.float64, copy=True)
    n = numpy.dot(q, q)
    if n < _EPS:
        return numpy.identity(4)
    q *= math.sqrt(2.0 / n)
    w, x, y, z = q
    return numpy.array([
        [1.0 - y * y - z * z, x * y - z * w, x * z + y * w, 0.0],
        [x * y + z * w, 1.0 - x * x - z * z, y * z - x * w, 0.0],
        [x * z - y * w, y * z + x * w, 1.0 - x * x - y * y, 0.0],
        [0.0, 0.0, 0.0, 1.0]])


def quaternion_from_matrix(matrix, isprecise=False):
    """Quaternion (w >= 0) of a rotation matrix.  isprecise=True: closed form with largest-diagonal
    pivoting; isprecise=False: principal eigenvector of the symmetric K matrix."""
    M = numpy.asarray(matrix, dtype=numpy.float64)[:4, :4]
    if isprecise:
        q = numpy.empty((4,))
        t = numpy.trace(M)
        if t > M[3, 3]:
            q[:] = (t, M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1])
        else:
            i, j, k = 0, 1, 2
            if M[1, 1] > M[0, 0]:
                i, j, k = 1, 2, 0
            if M[2, 2] > M[i, i]:
                i, j, k = 2, 0, 1
            t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
            v = numpy.empty((4,))
            v[i], v[j], v[k], v[3] = t, M[i, j] + M[j, i], M[k, i] + M[i, k], M[k, j] - M[j, k]
            q = v[[3, 0, 1, 2]]
        q = q * (0.5 / math.sqrt(t * M[3, 3]))
    else:
        (m00, m01, m02), (m10, m11, m12), (m20, m21, m22) = M[0, :3], M[1, :3], M[2, :3]
        K = numpy.array([[m00 - m11 - m22, 0.0, 0.0, 0.0],
                         [m01 + m10, m11 - m00 - m22, 0.0, 0.0],
                         [m02 + m20, m12 + m21, m22 - m00 - m11, 0.0],
                         [m21 - m12, m02 - m20, m10 - m01, m00 + m11 + m22]]) / 3.0
        w, V = numpy.linalg.eigh(K)
        q = V[[3, 0, 1, 2], numpy.argmax(w)]
    if q[0] < 0.0:
        q = -q
    return q
