"""Euler matrices and point-group symmetry lists (host-side helpers, numpy).

Python twin of csrc/host/symmetries.cpp; both follow xmippCore's conventions as
used by the reference path:
  * Euler_angles2matrix (ZYZ; pinned by src/xmipp/tests/test_binding.py:59-69 and
    the in-tree data/euler.cpp comparison, test_euler_main.cpp:26-56);
  * SymList::readSymmetryFile(name) + trueSymsNo()/getMatrices (RF.cpp:272-286):
    the list holds every group element except the identity
    (test_symmetries_main.cpp:44-51: i3h -> 119).
"""
import numpy as np


def euler_matrix(rot, tilt, psi):
    """A = Euler_angles2matrix(rot, tilt, psi), angles in degrees."""
    a, b, g = np.radians([rot, tilt, psi])
    ca, cb, cg = np.cos([a, b, g])
    sa, sb, sg = np.sin([a, b, g])
    cc, cs, sc, ss = cb * ca, cb * sa, sb * ca, sb * sa
    return np.array([[cg * cc - sg * sa, cg * cs + sg * ca, -cg * sb],
                     [-sg * cc - cg * sa, -sg * cs + cg * ca, sg * sb],
                     [sc, ss, cb]])


def _axis_rotation(axis, ang_deg):
    ax = np.asarray(axis, dtype=np.float64)
    ax = ax / np.linalg.norm(ax)
    t = np.radians(ang_deg)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * (K @ K)


def _mirror(normal):
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    return np.eye(3) - 2.0 * np.outer(n, n)


def _close_group(gens, limit=480):
    elems = [np.eye(3)]
    frontier = [np.eye(3)]
    while frontier:
        nxt = []
        for e in frontier:
            for g in gens:
                c = g @ e
                if not any(np.allclose(c, x, atol=1e-6) for x in elems):
                    elems.append(c)
                    nxt.append(c)
                    if len(elems) > limit:
                        raise ValueError("group does not close")
        frontier = nxt
    return elems


def _generators(name):
    s = name.strip().lower()
    rot = _axis_rotation
    if s in ("i", "i2", "i1", "i3", "i4", "ih", "i1h", "i2h", "i3h", "i4h"):
        h = s.endswith("h")
        base = s[:-1] if h else s
        if base in ("i", "i2"):
            g = [rot((0, 0, 1), 180), rot((0.525731114, 0, 0.850650807), 72), rot((0, 0.356822076, 0.934172364), 120)]
        elif base == "i1":
            g = [rot((1, 0, 0), 180), rot((0.85065080702670, 0, -0.5257311142635), 72), rot((0.9341723640, 0.3568220765, 0), 120)]
        elif base == "i3":
            g = [rot((-0.5257311143, 0, 0.8506508070), 180), rot((0, 0, 1), 72),
                 rot((-0.4911234778630044, 0.3568220764705179, 0.7946544753759428), 120)]
        else:  # i4
            g = [rot((0.5257311143, 0, 0.8506508070), 180), rot((0.8944271932547096, 0, 0.4472135909903704), 72),
                 rot((0.4911234778630044, 0.3568220764705179, 0.7946544753759428), 120)]
        if h:
            g.append(-np.eye(3))
        return g
    if s in ("t", "td", "th"):
        g = [rot((0, 0, 1), 120), rot((0, 0.816496, 0.577350), 180)]
        if s == "td":
            g.append(_mirror((1.4142136, 2.4494897, 0.0)))
        if s == "th":
            g.append(-np.eye(3))
        return g
    if s in ("o", "oh"):
        g = [rot((.5773502, .5773502, .5773502), 120), rot((0, 0, 1), 90)]
        if s == "oh":
            g.append(_mirror((0, 1, 1)))
        return g
    kind = s[0]
    tail = ""
    body = s[1:]
    while body and not body[-1].isdigit():
        tail = body[-1] + tail
        body = body[:-1]
    if kind not in ("c", "d", "s") or not body:
        raise ValueError("unknown symmetry group '%s'" % name)
    n = int(body)
    if kind == "s":
        if n % 2:
            raise ValueError("sN needs even N")
        return [rot((0, 0, 1), 360.0 / (n // 2)), -np.eye(3)]
    g = [rot((0, 0, 1), 360.0 / n)] if n > 1 else []
    if kind == "d":
        g.append(rot((1, 0, 0), 180))
    if tail == "v":
        g.append(_mirror((0, 1, 0) if kind == "c" else (1, 0, 0)))
    elif tail == "h":
        g.append(_mirror((0, 0, 1)))
    elif tail:
        raise ValueError("unknown symmetry group '%s'" % name)
    return g


def point_group_matrices(name):
    """All non-identity 3x3 matrices of the point group (what SL.getMatrices yields for
    isym in [0, trueSymsNo())), shape [n,3,3]."""
    gens = _generators(name)
    if not gens:
        return np.zeros((0, 3, 3))
    elems = _close_group(gens)
    return np.stack(elems[1:]) if len(elems) > 1 else np.zeros((0, 3, 3))
