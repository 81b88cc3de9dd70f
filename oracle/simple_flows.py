"""TEST INFRASTRUCTURE ONLY -- numpy oracle for the simple_flows boundary flavour
(SURVEY.md §8a row A9, Appendix A.3).

The reference for this flavour is itself numpy (no compiled path exists), so the
restatement is numpy too.  It is written as a *pull* from the pre-stream state,
the form the CUDA kernel uses, not as the reference's roll-then-overwrite
sequence; ``tests/test_oracle_golden.py`` proves both agree BITWISE against
golden vectors produced by executing the reference's own functions
(``tests/make_golden.py``).

Reference locations (relative to the upstream repository root):
  equilibrium       simulators/simple_flows/PoiseuilleFlow.py:25-42  (== slidingLid.py:33-50)
  moments           PoiseuilleFlow.py:49-53
  collision         PoiseuilleFlow.py:45-47   grid -= omega*(grid - feq)
  wall reflect      PoiseuilleFlow.py:59-74   (top/bottom, k in 1..nx-2 only)
  pressure columns  PoiseuilleFlow.py:76-88
  sliding lid       slidingLid.py:68-91       (four walls, rho_wall at the lid)
  step orders       PoiseuilleFlow.py:107-111 (Couette), :143-148 (Poiseuille), slidingLid.py:104-108
"""
import numpy as np

CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1])
CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1])


def feq(rho, ux, uy):
    """PoiseuilleFlow.py:25-42 -- operation order kept so results are bit-identical."""
    p3 = 3 * (ux + uy)
    m3 = 3 * (ux - uy)
    uu = 3 * (ux * ux + uy * uy)
    ux6 = 6 * ux
    uy6 = 6 * uy
    uxx9 = 9 * ux * ux
    uyy9 = 9 * uy * uy
    uxy9 = 9 * ux * uy
    a = 2 * rho / 9
    b = rho / 18
    c = rho / 36
    return np.array([a * (2 - uu),
                     b * (2 + ux6 + uxx9 - uu),
                     b * (2 + uy6 + uyy9 - uu),
                     b * (2 - ux6 + uxx9 - uu),
                     b * (2 - uy6 + uyy9 - uu),
                     c * (1 + p3 + uxy9 + uu),
                     c * (1 - m3 - uxy9 + uu),
                     c * (1 - p3 + uxy9 + uu),
                     c * (1 + m3 - uxy9 + uu)])


def moments(f):
    """PoiseuilleFlow.py:49-53 (np.sum over axis 0 accumulates f0..f8 in order)."""
    rho = ((((((((f[0] + f[1]) + f[2]) + f[3]) + f[4]) + f[5]) + f[6]) + f[7]) + f[8])
    ux = ((f[1] + f[5] + f[8]) - (f[3] + f[6] + f[7])) / rho
    uy = ((f[2] + f[5] + f[6]) - (f[4] + f[7] + f[8])) / rho
    return rho, ux, uy


def collide(f, omega):
    """PoiseuilleFlow.py:45-47 with the moments of :49-53; returns (rho, ux, uy)."""
    rho, ux, uy = moments(f)
    f -= omega * (f - feq(rho, ux, uy))
    return rho, ux, uy


def pull(f):
    """Periodic pull R[i](k,l) = f[i, k-cx, l-cy] (== the reference's np.roll stream)."""
    return np.stack([np.roll(f[i], (CX[i], CY[i]), axis=(0, 1)) for i in range(9)])


def reflect_top_bottom(post, R, uw, k_lo, k_hi, rho_wall=None):
    """PoiseuilleFlow.py:65-74 / slidingLid.py:83-91 as a gather from R.
    Wall layers are l=0 and l=T; fluid rows 1 and T-1 receive the reflected
    populations, for k in [k_lo, k_hi)."""
    T = R.shape[2] - 1
    k = slice(k_lo, k_hi)
    post[2, k, 1] = R[4, k, 0]
    post[5, k, 1] = R[7, k, 0]
    post[6, k, 1] = R[8, k, 0]
    post[4, k, T - 1] = R[2, k, T]
    if rho_wall is None:
        post[7, k, T - 1] = R[5, k, T] - 1 / 6 * uw
        post[8, k, T - 1] = R[6, k, T] + 1 / 6 * uw
    else:
        post[7, k, T - 1] = R[5, k, T] - 1 / 6 * uw * rho_wall
        post[8, k, T - 1] = R[6, k, T] + 1 / 6 * uw * rho_wall


def couette_step(f, omega, uw):
    """PoiseuilleFlow.py:107-111: moments -> collide -> stream -> reflect.
    Array (nx, ny+2), x periodic.  Returns the pre-collision ux (what the
    reference plots, :108,114)."""
    rho, ux, uy = collide(f, omega)
    R = pull(f)
    post = R.copy()
    reflect_top_bottom(post, R, uw, 1, f.shape[1] - 1)
    f[...] = post
    return ux


def pressure_columns(f, rho_in, rho_out):
    """PoiseuilleFlow.py:76-88: rewrite columns k=0 and k=X (all 9 channels, all l)."""
    rho, ux, uy = moments(f)
    e = feq(rho, ux, uy)
    X = f.shape[1] - 1
    e_in = feq(rho_in, ux[X - 1, :], uy[X - 1, :])
    f[:, 0, :] = e_in + (f[:, X - 1, :] - e[:, X - 1, :])
    e_out = feq(rho_out, ux[1, :], uy[1, :])
    f[:, X, :] = e_out + (f[:, 1, :] - e[:, 1, :])


def poiseuille_step(f, omega, rho_in, rho_out, uw=0.0):
    """PoiseuilleFlow.py:143-148.  Array (nx+2, ny+2).  Returns ux used in collide."""
    pressure_columns(f, rho_in, rho_out)
    R = pull(f)
    post = R.copy()
    reflect_top_bottom(post, R, uw, 1, f.shape[1] - 1)
    f[...] = post
    rho, ux, uy = collide(f, omega)
    return ux


def sliding_lid_step(f, omega, uw):
    """slidingLid.py:104-108 with bounce_back :68-91.  Array (L+2, L+2)."""
    R = pull(f)
    post = R.copy()
    X = f.shape[1] - 1
    T = f.shape[2] - 1
    li = slice(1, T)
    # left/right wall layers k=0, k=X -> fluid columns 1, X-1   (slidingLid.py:72-78)
    post[1, 1, li] = R[3, 0, li]
    post[5, 1, li] = R[7, 0, li]
    post[8, 1, li] = R[6, 0, li]
    post[3, X - 1, li] = R[1, X, li]
    post[6, X - 1, li] = R[8, X, li]
    post[7, X - 1, li] = R[5, X, li]
    # bottom and lid, k in 1..X-1; these overwrite the corner-adjacent cells written above
    k = slice(1, X)
    rho_wall = 2.0 * (R[2, k, T] + R[5, k, T] + R[6, k, T]) + R[0, k, T] + R[1, k, T] + R[3, k, T]
    reflect_top_bottom(post, R, uw, 1, X, rho_wall=rho_wall)
    f[...] = post
    rho, ux, uy = collide(f, omega)
    return ux


def sliding_lid_mpi_step(f, omega, uw):
    """slidingLidMPI.py:264-268 on one rank: stream, bounce_back_choosen (:180-204, all four walls, FULL index
    ranges, applied in place in the reference's order so that later lines read what earlier lines wrote),
    moments, collision.  Array (L+2, L+2)."""
    g = pull(f)
    g[3, -2, :] = g[1, -1, :]
    g[6, -2, :] = g[8, -1, :]
    g[7, -2, :] = g[5, -1, :]
    g[1, 1, :] = g[3, 0, :]
    g[5, 1, :] = g[7, 0, :]
    g[8, 1, :] = g[6, 0, :]
    g[2, :, 1] = g[4, :, 0]
    g[5, :, 1] = g[7, :, 0]
    g[6, :, 1] = g[8, :, 0]
    g[4, :, -2] = g[2, :, -1]
    g[7, :, -2] = g[5, :, -1] - 1 / 6 * uw
    g[8, :, -2] = g[6, :, -1] + 1 / 6 * uw
    f[...] = g
    collide(f, omega)


def obstacle_channel_step(f, omega, x0, x1, y0, y1, uw=0.0):
    """experimantal_flows/obstacle_canal.py:270-275 without the pressure step: stream, bounce_back_choosen with
    bottom / top walls (:320-329), apply_obstacle (:413-458, its ranges and channel pairs as written), moments,
    collision.  x is periodic."""
    g = pull(f)
    g[2, :, 1] = g[4, :, 0]
    g[5, :, 1] = g[7, :, 0]
    g[6, :, 1] = g[8, :, 0]
    g[4, :, -2] = g[2, :, -1]
    g[7, :, -2] = g[5, :, -1] - 1 / 6 * uw
    g[8, :, -2] = g[6, :, -1] + 1 / 6 * uw
    ys, xs = slice(y0, y1), slice(x0, x1)
    g[3, x0 - 1, ys] = g[1, x0, ys]
    g[7, x0 - 1, ys] = g[5, x0, ys]
    g[6, x0 - 1, ys] = g[8, x0, ys]
    g[1, x1 + 1, ys] = g[3, x1, ys]
    g[5, x1 + 1, ys] = g[7, x1, ys]
    g[6, x1 + 1, ys] = g[8, x1, ys]
    g[2, xs, y1 + 1] = g[4, xs, y1]
    g[5, xs, y1 + 1] = g[7, xs, y1]
    g[6, xs, y1 + 1] = g[8, xs, y1]
    g[4, xs, y0 - 1] = g[2, xs, y0]
    g[7, xs, y0 - 1] = g[5, xs, y0]
    g[8, xs, y0 - 1] = g[6, xs, y0]
    f[...] = g
    collide(f, omega)


def table_step(f, omega, cells, src, add):
    """The gather form the CUDA kernel applies (latticeboltzmann_b200/boundary_table.py): periodic pull, the
    listed cells from their tabulated pre-stream sources + constants, then collision."""
    g = pull(f).reshape(9, -1)
    vals = f.reshape(-1)[src]
    g[:, cells] = np.where(add != 0, vals + add, vals).T
    f[...] = g.reshape(f.shape)
    collide(f, omega)
