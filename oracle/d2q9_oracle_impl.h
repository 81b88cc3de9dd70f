/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the D2Q9 BGK time step.
 *
 * This header is included twice by d2q9_oracle.c, once with T=float (SUF=f32)
 * and once with T=double (SUF=f64). It restates, in plain C and without Eigen,
 * the arithmetic of the reference's opt2 path.  Every function cites the
 * reference file:line (relative to the upstream repository root) it follows.
 *
 * Nothing in the product path (latticeboltzmann_b200/) may call into this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Build flags mirror the reference build (setup.py:110-122 + CPython's default
 * CFLAGS): -O2, no -march, and -ffp-contract=off so that a*b+c is rounded
 * twice exactly as the x86-64 baseline build of the reference does.
 *
 * Parity status: the reference C++ cannot be compiled in this image (Eigen
 * 3.4.0 -- setup.py:32 -- is not vendored and not installed), so the one piece
 * of arithmetic that lives in Eigen rather than in the reference's own source,
 * the association order of `f_i.sum()` (c/d2q9.h:126), is restated from
 * Eigen 3.4.0's published algorithm (Core/Redux.h, redux_novec_unroller:
 * recursive halving, Length 9 -> 4 + 5 -> (2+2) + (2 + (1+2))).  Everything
 * else is pinned: streaming/boundaries bit-exactly against the reference's own
 * Python functions, equilibrium/collide against the reference's numpy
 * formulas (tests/02-CollideTest.py:43-89, cavity_opt1.py:100-160) -- see
 * tests/make_golden.py and tests/test_oracle_golden.py.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* c/d2q9.h:59-81 -- d2q9_equilibrium1<T>.  Expression order is kept exactly:
 * C evaluates `1 + ux + ux*ux/2 - uu` as ((1 + ux) + (ux*ux)/2) - uu.       */
static inline void FN(eq1)(T rho, T ux, T uy, T *out)
{
    T w_0 = 4 * rho / 9;      /* (4*rho)/9                      d2q9.h:61 */
    T w_1234 = rho / 9;       /*                                d2q9.h:62 */
    T w_5678 = rho / 36;      /*                                d2q9.h:63 */
    ux *= 3;                  /*                                d2q9.h:64 */
    uy *= 3;                  /*                                d2q9.h:65 */
    T cu5 = ux + uy;          /*                                d2q9.h:66 */
    T cu6 = -ux + uy;         /*                                d2q9.h:67 */
    T cu7 = -ux - uy;         /*                                d2q9.h:68 */
    T cu8 = ux - uy;          /*                                d2q9.h:69 */
    T uu = (ux * ux + uy * uy) / 6;                          /* d2q9.h:70 */
    out[0] = w_0 * (1 - uu);                                 /* d2q9.h:72 */
    out[1] = w_1234 * (1 + ux + ux * ux / 2 - uu);           /* d2q9.h:73 */
    out[2] = w_1234 * (1 + uy + uy * uy / 2 - uu);           /* d2q9.h:74 */
    out[3] = w_1234 * (1 - ux + ux * ux / 2 - uu);           /* d2q9.h:75 */
    out[4] = w_1234 * (1 - uy + uy * uy / 2 - uu);           /* d2q9.h:76 */
    out[5] = w_5678 * (1 + cu5 + cu5 * cu5 / 2 - uu);        /* d2q9.h:77 */
    out[6] = w_5678 * (1 + cu6 + cu6 * cu6 / 2 - uu);        /* d2q9.h:78 */
    out[7] = w_5678 * (1 + cu7 + cu7 * cu7 / 2 - uu);        /* d2q9.h:79 */
    out[8] = w_5678 * (1 + cu8 + cu8 * cu8 / 2 - uu);        /* d2q9.h:80 */
}

void FN(orc_equilibrium1)(T rho, T ux, T uy, T *out9) { FN(eq1)(rho, ux, uy, out9); }

/* c/d2q9.h:98-108 -- d2q9_equilibriumn<T>: f is (9, n) row-major, column kl
 * is addressed with stride n (Eigen::Map with Stride(1, cols), d2q9.h:105). */
void FN(orc_equilibriumn)(const T *rho, const T *ux, const T *uy, T *f, int64_t n)
{
    for (int64_t kl = 0; kl < n; ++kl) {
        T e[9];
        FN(eq1)(rho[kl], ux[kl], uy[kl], e);
        for (int i = 0; i < 9; ++i) f[i * n + kl] = e[i];
    }
}

/* One cell of c/d2q9.h:125-129.  rho follows Eigen 3.4.0's unrolled
 * non-vectorised redux tree (see the file header).                          */
static inline void FN(collide1)(T *f, int64_t stride, T omega)
{
    T f0 = f[0], f1 = f[stride], f2 = f[2 * stride], f3 = f[3 * stride], f4 = f[4 * stride];
    T f5 = f[5 * stride], f6 = f[6 * stride], f7 = f[7 * stride], f8 = f[8 * stride];
    T rho = ((f0 + f1) + (f2 + f3)) + ((f4 + f5) + (f6 + (f7 + f8)));  /* d2q9.h:126 */
    T ux = (f1 - f3 + f5 - f6 - f7 + f8) / rho;                       /* d2q9.h:127 */
    T uy = (f2 - f4 + f5 + f6 - f7 - f8) / rho;                       /* d2q9.h:128 */
    T e[9];
    FN(eq1)(rho, ux, uy, e);
    f[0] = f0 + omega * (e[0] - f0);                                  /* d2q9.h:129 */
    f[stride] = f1 + omega * (e[1] - f1);
    f[2 * stride] = f2 + omega * (e[2] - f2);
    f[3 * stride] = f3 + omega * (e[3] - f3);
    f[4 * stride] = f4 + omega * (e[4] - f4);
    f[5 * stride] = f5 + omega * (e[5] - f5);
    f[6 * stride] = f6 + omega * (e[6] - f6);
    f[7 * stride] = f7 + omega * (e[7] - f7);
    f[8 * stride] = f8 + omega * (e[8] - f8);
}

/* c/d2q9.h:121-131 -- d2q9_colliden<T>, in place on a (9, n) row-major array. */
void FN(orc_colliden)(T *f, int64_t n, T omega)
{
    for (int64_t kl = 0; kl < n; ++kl) FN(collide1)(f + kl, n, omega);
}

/* Same loop split over [lo, hi) so the CPU baseline can run one slab per thread
 * or process (the reference's parallelism is one MPI rank per block).        */
void FN(orc_colliden_range)(T *f, int64_t n, T omega, int64_t lo, int64_t hi)
{
    for (int64_t kl = lo; kl < hi; ++kl) FN(collide1)(f + kl, n, omega);
}

/* PyLB/Streaming.py:33-46 -- stream(): for i = 1..8
 *   f[i] = np.roll(f[i], c_ic[i], axis=(0, 1))
 * i.e. out[k][l] = in[(k - cx) mod nx][(l - cy) mod ny].  `tmp` must hold
 * nx*ny values (np.roll allocates a fresh array per channel as well).        */
static const int FN(CX)[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};   /* Streaming.py:28 */
static const int FN(CY)[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};   /* Streaming.py:29 */

void FN(orc_stream)(T *f, int64_t nx, int64_t ny, T *tmp)
{
    int64_t n = nx * ny;
    for (int i = 1; i < 9; ++i) {
        T *fi = f + i * n;
        int cx = FN(CX)[i], cy = FN(CY)[i];
        for (int64_t k = 0; k < nx; ++k) {
            int64_t ks = (k - cx + nx) % nx;
            for (int64_t l = 0; l < ny; ++l) {
                int64_t ls = (l - cy + ny) % ny;
                tmp[k * ny + l] = fi[ks * ny + ls];
            }
        }
        memcpy(fi, tmp, (size_t)n * sizeof(T));
    }
}

/* simulators/parallel_lid_drive_cavity/cavity_opt2.py:109-177 --
 * stream_and_bounce_back(f_ikl, u0), followed literally: snapshot the four
 * edges, roll, then the ordered overwrites.  Channel names (cavity_opt2.py:70):
 * E=1 N=2 W=3 S=4 NE=5 NW=6 SW=7 SE=8.  `walls_lr` = 1 is the shipped code
 * path (`if True:` at :148); 0 is the documented Couette variant (:147).
 * scratch must hold nx*ny + 2*9*nx + 2*9*ny values.                          */
void FN(orc_cavity_stream_and_bounce_back)(T *f, int64_t nx, int64_t ny, T u0, int walls_lr, T *scratch)
{
    const int E = 1, N = 2, W = 3, S = 4, NE = 5, NW = 6, SW = 7, SE = 8;
    int64_t n = nx * ny;
    T *tmp = scratch;
    T *fbottom = tmp + n;          /* (9, nx)  f[:, :, 0]    :125 */
    T *ftop = fbottom + 9 * nx;    /* (9, nx)  f[:, :, -1]   :126 */
    T *fleft = ftop + 9 * nx;      /* (9, ny)  f[:, 0, :]    :128 */
    T *fright = fleft + 9 * ny;    /* (9, ny)  f[:, -1, :]   :129 */
#define F(i, k, l) f[(i) * n + (k) * ny + (l)]
    for (int i = 0; i < 9; ++i) {
        for (int64_t k = 0; k < nx; ++k) {
            fbottom[i * nx + k] = F(i, k, 0);
            ftop[i * nx + k] = F(i, k, ny - 1);
        }
        for (int64_t l = 0; l < ny; ++l) {
            fleft[i * ny + l] = F(i, 0, l);
            fright[i * ny + l] = F(i, nx - 1, l);
        }
    }
    FN(orc_stream)(f, nx, ny, tmp);                                           /* :131 */

    for (int64_t k = 0; k < nx; ++k) {                                        /* :134-136 */
        F(N, k, 0) = fbottom[S * nx + k];
        F(NE, k, 0) = fbottom[SW * nx + k];
        F(NW, k, 0) = fbottom[SE * nx + k];
    }
    /* w_i = np.array([... 1/36 ...], dtype=dtype) (:90): 1/36 is computed in
     * double and rounded to T; 6*w_i[D.SE] is then rounded in T (numpy scalar
     * times weak Python int), multiplied by rho (T), then by u0 cast to T.   */
    const T w8 = (T)(1.0 / 36.0);
    const T six_w = (T)6 * w8;
    for (int64_t k = 0; k < nx; ++k) {                                        /* :140-145 */
        T rho = ftop[NW * nx + k] + ftop[N * nx + k] + ftop[NE * nx + k]
              + F(NW, k, ny - 1) + F(N, k, ny - 1) + F(NE, k, ny - 1)
              + F(W, k, ny - 1) + F(0, k, ny - 1) + F(E, k, ny - 1);
        F(S, k, ny - 1) = ftop[N * nx + k];
        F(SE, k, ny - 1) = ftop[NW * nx + k] + six_w * rho * u0;
        F(SW, k, ny - 1) = ftop[NE * nx + k] - six_w * rho * u0;
    }
    if (walls_lr) {                                                           /* :148 */
        for (int64_t l = 0; l < ny; ++l) {                                    /* :150-157 */
            F(E, 0, l) = fleft[W * ny + l];
            F(NE, 0, l) = fleft[SW * ny + l];
            F(SE, 0, l) = fleft[NW * ny + l];
            F(W, nx - 1, l) = fright[E * ny + l];
            F(NW, nx - 1, l) = fright[SE * ny + l];
            F(SW, nx - 1, l) = fright[NE * ny + l];
        }
        F(N, 0, 0) = fbottom[S * nx + 0];                                     /* :160-162 */
        F(E, 0, 0) = fbottom[W * nx + 0];
        F(NE, 0, 0) = fbottom[SW * nx + 0];
        F(N, nx - 1, 0) = fbottom[S * nx + nx - 1];                           /* :165-167 */
        F(W, nx - 1, 0) = fbottom[E * nx + nx - 1];
        F(NW, nx - 1, 0) = fbottom[SE * nx + nx - 1];
        F(S, 0, ny - 1) = ftop[N * nx + 0];                                   /* :170-172 */
        F(E, 0, ny - 1) = ftop[W * nx + 0];
        F(SE, 0, ny - 1) = ftop[NW * nx + 0];
        F(S, nx - 1, ny - 1) = ftop[N * nx + nx - 1];                         /* :175-177 */
        F(W, nx - 1, ny - 1) = ftop[E * nx + nx - 1];
        F(SW, nx - 1, ny - 1) = ftop[NE * nx + nx - 1];
    }
#undef F
}

/* cavity_opt2.py:272-277 on a single rank (communicate() is a no-op there:
 * every neighbour is MPI.PROC_NULL): nsteps x { stream_and_bounce_back;
 * collide }.                                                                 */
void FN(orc_cavity_run)(T *f, int64_t nx, int64_t ny, T omega, T u0, int walls_lr, int64_t nsteps, T *scratch)
{
    for (int64_t s = 0; s < nsteps; ++s) {
        FN(orc_cavity_stream_and_bounce_back)(f, nx, ny, u0, walls_lr, scratch);
        FN(orc_colliden)(f, nx * ny, omega);
    }
}

/* simulators/serial_shear_wave/Python/shear_wave_opt2.py:95-99 --
 * nsteps x { stream; collide; amplitude }.  ampl (may be NULL) receives
 *   sum_k uy(k, ny//2) * uy_k[k] * 2/nx
 * with uy = (f2 - f4 + f5 + f6 - f7 - f8)/rho in plain left-to-right order
 * (numpy's dot/sum orders are not specified; tests compare with a tolerance). */
void FN(orc_periodic_run)(T *f, int64_t nx, int64_t ny, T omega, int64_t nsteps, const T *uy_k, T *ampl, T *scratch)
{
    int64_t n = nx * ny;
    for (int64_t s = 0; s < nsteps; ++s) {
        FN(orc_stream)(f, nx, ny, scratch);
        FN(orc_colliden)(f, n, omega);
        if (ampl) {
            int64_t l = ny / 2;
            T acc = 0;
            for (int64_t k = 0; k < nx; ++k) {
                const T *c = f + k * ny + l;
                T rho = c[0] + c[n] + c[2 * n] + c[3 * n] + c[4 * n] + c[5 * n] + c[6 * n] + c[7 * n] + c[8 * n];
                T uy = (c[2 * n] - c[4 * n] + c[5 * n] + c[6 * n] - c[7 * n] - c[8 * n]) / rho;
                acc += uy * uy_k[k];
            }
            ampl[s] = acc * 2 / (T)nx;
        }
    }
}

/* The same cavity step written as one fused pull loop over a second array
 * (derived from cavity_opt2.py:109-177 and verified
 * bitwise against orc_cavity_stream_and_bounce_back by tests/test_oracle_*).
 * Used only as the "CPU-fused" informational baseline and as an independent
 * cross-check of the pull formulation the CUDA kernel uses.
 * rows [k_lo, k_hi) of dst are produced from src (both (9, nx, ny)).         */
static void FN(orc_cavity_step_pull_rows_cols)(const T *src, T *dst, int64_t nx, int64_t ny, T omega, T u0,
                                               int walls_lr, int do_collide, int64_t k, int64_t l_lo, int64_t l_hi)
{
    static const int OPP[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    const int E = 1, N = 2, W = 3, NE = 5, NW = 6, SW = 7, SE = 8;
    int64_t n = nx * ny, X = nx - 1, Tt = ny - 1;
    const T six_w = (T)6 * (T)(1.0 / 36.0);
#define SRC(i, k, l) src[(i) * n + (k) * ny + (l)]
    {
        for (int64_t l = l_lo; l < l_hi; ++l) {
            T R[9], p[9];
            for (int i = 0; i < 9; ++i) {
                int cx = FN(CX)[i], cy = FN(CY)[i];
                R[i] = SRC(i, (k - cx + nx) % nx, (l - cy + ny) % ny);
                int outside = (cy == 1 && l == 0) || (cy == -1 && l == Tt);
                if (walls_lr) outside = outside || (cx == 1 && k == 0) || (cx == -1 && k == X);
                p[i] = outside ? SRC(OPP[i], k, l) : R[i];
            }
            if (l == Tt) {
                T rho = SRC(NW, k, l) + SRC(N, k, l) + SRC(NE, k, l) + R[NW] + R[N] + R[NE] + R[W] + R[0] + R[E];
                if (!walls_lr || k >= 1) p[SE] = SRC(NW, k, l) + six_w * rho * u0;
                if (!walls_lr || k <= X - 1) p[SW] = SRC(NE, k, l) - six_w * rho * u0;
            }
            if (do_collide) FN(collide1)(p, 1, omega);
            for (int i = 0; i < 9; ++i) dst[i * n + k * ny + l] = p[i];
        }
    }
#undef SRC
}

void FN(orc_cavity_step_pull_rows)(const T *src, T *dst, int64_t nx, int64_t ny, T omega, T u0,
                                   int walls_lr, int do_collide, int64_t k_lo, int64_t k_hi)
{
    for (int64_t k = k_lo; k < k_hi; ++k)
        FN(orc_cavity_step_pull_rows_cols)(src, dst, nx, ny, omega, u0, walls_lr, do_collide, k, 0, ny);
}

/* orc_cavity_step_pull_rows with the index arithmetic hoisted out of the column loop: the same per-cell rule
 * (the literal function above is the definition; tests/test_oracle_golden.py compares the two bitwise), but
 * columns 1 .. ny-2 -- which never see the top/bottom wall, the lid or the y wrap -- read their nine sources
 * through row pointers instead of two modulo operations per population.  This is what makes the oracle usable
 * at BASELINE's 4096^2 (tests/test_gpu_parity.py) and on the benchmarked strips (bench.py cpu_baseline leg). */
void FN(orc_cavity_step_pull_rows_hoisted)(const T *src, T *dst, int64_t nx, int64_t ny, T omega, T u0,
                                           int walls_lr, int do_collide, int64_t k_lo, int64_t k_hi)
{
    static const int OPP[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    int64_t n = nx * ny, X = nx - 1;
    for (int64_t k = k_lo; k < k_hi; ++k) {
        /* the two wall columns of this row: the literal rule */
        FN(orc_cavity_step_pull_rows_cols)(src, dst, nx, ny, omega, u0, walls_lr, do_collide, k, 0, 1);
        if (ny > 1) FN(orc_cavity_step_pull_rows_cols)(src, dst, nx, ny, omega, u0, walls_lr, do_collide, k, ny - 1, ny);
        const T *row[9];
        int bounce[9];
        for (int i = 0; i < 9; ++i) {
            int cx = FN(CX)[i], cy = FN(CY)[i];
            bounce[i] = walls_lr && ((cx == 1 && k == 0) || (cx == -1 && k == X));
            /* source row pointer, already shifted by -cy: element l of it is src[i, k - cx, l - cy] */
            row[i] = bounce[i] ? src + OPP[i] * n + k * ny : src + i * n + ((k - cx + nx) % nx) * ny - cy;
        }
        for (int64_t l = 1; l < ny - 1; ++l) {
            T p[9];
            for (int i = 0; i < 9; ++i) p[i] = row[i][l];
            if (do_collide) FN(collide1)(p, 1, omega);
            for (int i = 0; i < 9; ++i) dst[i * n + k * ny + l] = p[i];
        }
    }
}

/* Fully periodic fused pull step. */
void FN(orc_periodic_step_pull_rows)(const T *src, T *dst, int64_t nx, int64_t ny, T omega,
                                     int do_collide, int64_t k_lo, int64_t k_hi)
{
    int64_t n = nx * ny;
    for (int64_t k = k_lo; k < k_hi; ++k) {
        for (int64_t l = 0; l < ny; ++l) {
            T p[9];
            for (int i = 0; i < 9; ++i)
                p[i] = src[i * n + ((k - FN(CX)[i] + nx) % nx) * ny + ((l - FN(CY)[i] + ny) % ny)];
            if (do_collide) FN(collide1)(p, 1, omega);
            for (int i = 0; i < 9; ++i) dst[i * n + k * ny + l] = p[i];
        }
    }
}

/* cavity_opt2.py:280-281 -- rho = sum_i f_i (numpy sums axis 0 sequentially
 * for a C-contiguous (9, nx, ny) array: ((f0+f1)+f2)+...), u = (f^T . c)/rho.
 * The dot's internal order is BLAS-defined; tests use a tolerance for u.     */
void FN(orc_moments)(const T *f, int64_t n, T *rho, T *ux, T *uy)
{
    for (int64_t kl = 0; kl < n; ++kl) {
        const T *c = f + kl;
        T r = c[0] + c[n] + c[2 * n] + c[3 * n] + c[4 * n] + c[5 * n] + c[6 * n] + c[7 * n] + c[8 * n];
        rho[kl] = r;
        ux[kl] = (c[n] - c[3 * n] + c[5 * n] - c[6 * n] - c[7 * n] + c[8 * n]) / r;
        uy[kl] = (c[2 * n] - c[4 * n] + c[5 * n] + c[6 * n] - c[7 * n] - c[8 * n]) / r;
    }
}

#undef FN
#undef CAT
#undef CAT_
