"""Oracle rigid-water constraints (numpy float64).  TEST INFRASTRUCTURE ONLY.

The reference's water drivers build the system with ``constrained=True``
(code/water/test_script/test_nosehoover.py:33-37) and its integrator programs call
``addConstrainPositions`` / ``addConstrainVelocities`` (code/hack_integrator.py:146-165, :274-277, :421-422); the
arithmetic lives in OpenMM (SETTLE for 3-site water), which is not installed here: parity with OpenMM's
implementation is UNPINNED.  Restated here:

  settle_positions   the analytic SETTLE of Miyamoto & Kollman, J. Comput. Chem. 13, 952 (1992): given positions x0
                     that satisfy the constraints and unconstrained positions x1, the constrained positions
                     x1 + sum_j lambda_ij (x0_i - x0_j) / m_i  (displacements along the OLD bond vectors)
  shake_positions    the same equations solved by SHAKE iteration to 1e-14 - an independent check of the closed form
  settle_velocities  removes the relative velocity along the three bonds with impulses along them (RATTLE velocity
                     stage for a rigid triangle, a 3x3 linear solve - the closed form of M&K appendix B)

Molecules are consecutive triplets [O, H, H]; distances: d_oh (O-H), d_hh (H-H).
TIP3P geometry (OpenMM tip3p.xml [3P-memory]): O-H 0.09572 nm, H-O-H 104.52 deg.
"""
import numpy as np

TIP3P_OH = 0.09572
TIP3P_HH = 2.0 * 0.09572 * np.sin(np.deg2rad(104.52) / 2.0)


def settle_positions(x0, x1, masses, d_oh=TIP3P_OH, d_hh=TIP3P_HH):
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1, 3, 3)
    x1 = np.asarray(x1, dtype=np.float64).reshape(-1, 3, 3)
    m = np.asarray(masses, dtype=np.float64).reshape(-1, 3)
    out = np.empty_like(x1)
    for k in range(x0.shape[0]):
        out[k] = _settle_one(x0[k], x1[k], m[k], d_oh, d_hh)
    return out.reshape(-1, 3)


def _settle_one(a0, a1, m, d_oh, d_hh):
    m0, m1, m2 = m
    xp0, xp1, xp2 = a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]      # displacements of the unconstrained step
    xb0, xc0 = a0[1] - a0[0], a0[2] - a0[0]
    inv_mt = 1.0 / (m0 + m1 + m2)
    xcom = (xp0 * m0 + (xb0 + xp1) * m1 + (xc0 + xp2) * m2) * inv_mt
    xa1 = xp0 - xcom
    xb1 = xb0 + xp1 - xcom
    xc1 = xc0 + xp2 - xcom
    akz = np.cross(xb0, xc0)
    akx = np.cross(xa1, akz)
    aky = np.cross(akz, akx)
    t1, t2, t3 = akx / np.linalg.norm(akx), aky / np.linalg.norm(aky), akz / np.linalg.norm(akz)
    xb0d, yb0d = t1 @ xb0, t2 @ xb0
    xc0d, yc0d = t1 @ xc0, t2 @ xc0
    za1d = t3 @ xa1
    xb1d, yb1d, zb1d = t1 @ xb1, t2 @ xb1, t3 @ xb1
    xc1d, yc1d, zc1d = t1 @ xc1, t2 @ xc1, t3 @ xc1
    rc = 0.5 * d_hh
    rb = np.sqrt(d_oh * d_oh - rc * rc)
    ra = rb * (m1 + m2) * inv_mt
    rb -= ra
    sinphi = za1d / ra
    cosphi = np.sqrt(1 - sinphi * sinphi)
    sinpsi = (zb1d - zc1d) / (2 * rc * cosphi)
    cospsi = np.sqrt(1 - sinpsi * sinpsi)
    ya2d = ra * cosphi
    xb2d = -rc * cospsi
    yb2d = -rb * cosphi - rc * sinpsi * sinphi
    yc2d = -rb * cosphi + rc * sinpsi * sinphi
    xb2d2 = xb2d * xb2d
    hh2 = 4.0 * xb2d2 + (yb2d - yc2d) ** 2 + (zb1d - zc1d) ** 2
    deltx = 2.0 * xb2d + np.sqrt(4.0 * xb2d2 - hh2 + d_hh * d_hh)
    xb2d -= deltx * 0.5
    alpha = xb2d * (xb0d - xc0d) + yb0d * yb2d + yc0d * yc2d
    beta = xb2d * (yc0d - yb0d) + xb0d * yb2d + xc0d * yc2d
    gamma = xb0d * yb1d - xb1d * yb0d + xc0d * yc1d - xc1d * yc0d
    al2be2 = alpha * alpha + beta * beta
    sintheta = (alpha * gamma - beta * np.sqrt(al2be2 - gamma * gamma)) / al2be2
    costheta = np.sqrt(1 - sintheta * sintheta)
    xa3d, ya3d, za3d = -ya2d * sintheta, ya2d * costheta, za1d
    xb3d, yb3d, zb3d = xb2d * costheta - yb2d * sintheta, xb2d * sintheta + yb2d * costheta, zb1d
    xc3d, yc3d, zc3d = -xb2d * costheta - yc2d * sintheta, -xb2d * sintheta + yc2d * costheta, zc1d
    xa3 = t1 * xa3d + t2 * ya3d + t3 * za3d
    xb3 = t1 * xb3d + t2 * yb3d + t3 * zb3d
    xc3 = t1 * xc3d + t2 * yc3d + t3 * zc3d
    return np.stack([a0[0] + xcom + xa3, a0[1] + xcom + xb3 - xb0, a0[2] + xcom + xc3 - xc0])


def shake_positions(x0, x1, masses, d_oh=TIP3P_OH, d_hh=TIP3P_HH, tol=1e-14, max_iter=2000):
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1, 3, 3)
    x = np.array(x1, dtype=np.float64).reshape(-1, 3, 3)
    m = np.asarray(masses, dtype=np.float64).reshape(-1, 3)
    pairs = ((0, 1, d_oh), (0, 2, d_oh), (1, 2, d_hh))
    for k in range(x.shape[0]):
        for _ in range(max_iter):
            worst = 0.0
            for i, j, d in pairs:
                r = x[k, i] - x[k, j]
                diff = r @ r - d * d
                worst = max(worst, abs(diff) / (d * d))
                r0 = x0[k, i] - x0[k, j]
                g = diff / (2.0 * (r @ r0) * (1.0 / m[k, i] + 1.0 / m[k, j]))
                x[k, i] -= g * r0 / m[k, i]
                x[k, j] += g * r0 / m[k, j]
            if worst < tol:
                break
    return x.reshape(-1, 3)


def settle_velocities(x, v, masses):
    x = np.asarray(x, dtype=np.float64).reshape(-1, 3, 3)
    v = np.array(v, dtype=np.float64).reshape(-1, 3, 3)
    m = np.asarray(masses, dtype=np.float64).reshape(-1, 3)
    for k in range(x.shape[0]):
        a, b, c = x[k]
        ma, mb, mc = m[k]
        eab, ebc, eca = b - a, c - b, a - c
        eab, ebc, eca = eab / np.linalg.norm(eab), ebc / np.linalg.norm(ebc), eca / np.linalg.norm(eca)
        vab, vbc, vca = (v[k, 1] - v[k, 0]) @ eab, (v[k, 2] - v[k, 1]) @ ebc, (v[k, 0] - v[k, 2]) @ eca
        # v_a += (t_ab e_ab - t_ca e_ca)/m_a, v_b += (t_bc e_bc - t_ab e_ab)/m_b, v_c += (t_ca e_ca - t_bc e_bc)/m_c
        # so that the relative velocity along each bond vanishes
        A = np.array([[-(1 / ma + 1 / mb), (ebc @ eab) / mb, (eca @ eab) / ma],
                      [(eab @ ebc) / mb, -(1 / mb + 1 / mc), (eca @ ebc) / mc],
                      [(eab @ eca) / ma, (ebc @ eca) / mc, -(1 / mc + 1 / ma)]])
        tab, tbc, tca = np.linalg.solve(A, -np.array([vab, vbc, vca]))
        v[k, 0] += (tab * eab - tca * eca) / ma
        v[k, 1] += (tbc * ebc - tab * eab) / mb
        v[k, 2] += (tca * eca - tbc * ebc) / mc
    return v.reshape(-1, 3)


def rigid_water(n_mol, rng, box=2.0, d_oh=TIP3P_OH, d_hh=TIP3P_HH):
    """n_mol rigid [O,H,H] molecules with random positions / orientations (nm)."""
    half = np.arcsin(0.5 * d_hh / d_oh)
    local = np.array([[0, 0, 0], [d_oh * np.sin(half), 0, d_oh * np.cos(half)], [-d_oh * np.sin(half), 0, d_oh * np.cos(half)]])
    q = rng.standard_normal((n_mol, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    c = rng.uniform(0, box, (n_mol, 1, 3))
    return (c + np.einsum("mij,sj->msi", R, local)).reshape(-1, 3)
