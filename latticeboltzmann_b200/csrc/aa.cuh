// In-place streaming ("AA pattern", Bailey et al. 2009): ONE copy of the populations instead of the A/B pair --
// 32768^2 fp64 in 77 GB instead of 155 GB -- at the same 144 B per cell per step.
//
// The stored state is what the A/B kernels store: the post-collision populations f*_i(x) ("pre" of the next
// step's pull).  Two kinds of step alternate, both in place and race-free because every thread
// writes exactly the locations it read:
//
//   EVEN step (natural layout -> swapped layout).  Cell x gathers post[i] = pre[i](x - c_i) from its NEIGHBOURS'
//     slots i (the same pull as step_kernel), or -- if that source lies outside a wall -- its own slot opp(i)
//     (half-way bounce-back, cavity_opt2.py:134-177); applies the lid; collides; and stores the new f*_opp(i)(x)
//     back into the very location post[i] came from.  Afterwards f*_j(x) sits at (x + c_j, slot opp(j)), or at
//     (x, slot j) where x + c_j is outside a wall.
//   ODD step (swapped -> natural).  By construction the value cell z needs for direction i -- f*_i(z - c_i), or
//     f*_opp(i)(z) for a bounced direction -- now sits in z's OWN slot opp(i): the pull is nine local, aligned
//     loads; the results go back to the natural slots.
//
// One block whose periodic rings close on itself; the wrap is index arithmetic (no ghost frame).  The lid's
// density (cavity_opt2.py:140-142) reads periodically rolled values, which at the two top corners of a walled box
// belong to the OPPOSITE corner cell (the reference's wrap quirk) -- locations another thread
// rewrites in the same step -- so a four-value stash is taken by a tiny kernel before each step.
// Arithmetic and rule order are step_kernel.cuh's: results are bit-identical to the A/B kernels.
#pragma once
#include "step_kernel.cuh"

namespace lbm {

template <int BC>
__device__ __forceinline__ bool aa_outside(const int i, int k, int l, int lnx, int lny)
{
    if (BC == BC_PERIODIC) return false;
    const bool walls = BC == BC_CAVITY;
    return (cy_of(i) == 1 && l == 0) || (cy_of(i) == -1 && l == lny - 1) || (walls && cx_of(i) == 1 && k == 0) ||
           (walls && cx_of(i) == -1 && k == lnx - 1);
}
// The four periodically rolled values the lid reads across the box at the top corners of a walled cavity:
// stash[0] = pre[E](X, T), stash[1] = pre[NE](X, T-1)  (for cell (0, T));  stash[2] = pre[W](0, T), stash[3] = pre[NW](0, T-1).
template <typename T>
__global__ void aa_corner_stash_kernel(const __grid_constant__ StepParams<T> p, T *stash)
{
    const int X = p.lnx - 1, Tt = p.lny - 1;
    const T *buf = p.buf[0];
    if (threadIdx.x == 0) stash[0] = aa_natural<T>(p, buf, QE, X, Tt);
    if (threadIdx.x == 1) stash[1] = aa_natural<T>(p, buf, QNE, X, wrap_idx(Tt - 1, p.lny));
    if (threadIdx.x == 2) stash[2] = aa_natural<T>(p, buf, QW, 0, Tt);
    if (threadIdx.x == 3) stash[3] = aa_natural<T>(p, buf, QNW, 0, wrap_idx(Tt - 1, p.lny));
}

// Lid rule on the gathered populations f (cavity_opt2.py:140-145 as a gather; same operations as wall_rules).
// On the top row f[SE], f[S], f[SW] hold the cell's own pre[NW], pre[N], pre[NE] (bounced), the others R[.].
template <typename T, int BC>
__device__ __forceinline__ void aa_lid(const StepParams<T> &p, const T *stash, T (&f)[9], int k, int l)
{
    if (BC == BC_PERIODIC || l != p.lny - 1) return;
    const bool walls = BC == BC_CAVITY;
    const bool left = walls && k == 0, right = walls && k == p.lnx - 1;
    const T o_nw = f[QSE], o_n = f[QS], o_ne = f[QSW];
    const T r_nw = right ? stash[3] : f[QNW], r_ne = left ? stash[1] : f[QNE];
    const T r_w = right ? stash[2] : f[QW], r_e = left ? stash[0] : f[QE];
    T rho = rn_add(o_nw, o_n);
    rho = rn_add(rho, o_ne);
    rho = rn_add(rho, r_nw);
    rho = rn_add(rho, f[QN]);
    rho = rn_add(rho, r_ne);
    rho = rn_add(rho, r_w);
    rho = rn_add(rho, f[Q0]);
    rho = rn_add(rho, r_e);
    const T six_w = rn_mul(T(6), T(1.0 / 36.0));
    const T lid = rn_mul(rn_mul(six_w, rho), p.u_wall);
    if (!left) f[QSE] = rn_add(o_nw, lid);
    if (!right) f[QSW] = rn_sub(o_ne, lid);
}

template <typename T, int BC, bool EXACT, bool ODD>
__global__ void __launch_bounds__(TILE_L, min_ctas_per_sm<T, EXACT>()) aa_step_kernel(const __grid_constant__ StepParams<T> p, const T *__restrict__ stash)
{
    T *buf = p.buf[0];
    const int kt = (int)blockIdx.x / p.tiles_l, lt = (int)blockIdx.x - kt * p.tiles_l;
    const int l = lt * TILE_L + threadIdx.x;
    if (l >= p.lny) return;
    const int k0 = kt * p.rows_per_tile, k1 = min(k0 + p.rows_per_tile, p.lnx);
    const bool l_rim = l == 0 || l == p.lny - 1;
    for (int k = k0; k < k1; ++k) {
        T *own = buf + (long long)(k + 1) * p.pitch + (l + PAD_L);
        T f[9];
        if (ODD) {
            // swapped -> natural: everything this cell needs is in its own slots
#pragma unroll
            for (int i = 0; i < 9; ++i) f[i] = own[(long long)opp_of(i) * p.pop_stride];
            aa_lid<T, BC>(p, stash, f, k, l);
            d2q9_collide<T, EXACT>(f, p.omega);
#pragma unroll
            for (int i = 0; i < 9; ++i) own[(long long)i * p.pop_stride] = f[i];
        } else if (!l_rim && k > 0 && k < p.lnx - 1) {
            // natural -> swapped, interior: the pull of step_kernel, results stored where they were read
            char *sp = reinterpret_cast<char *>(own);
#pragma unroll
            for (int i = 0; i < 9; ++i) f[i] = *reinterpret_cast<const T *>(sp + p.ld_off[i]);
            d2q9_collide<T, EXACT>(f, p.omega);
#pragma unroll
            for (int i = 0; i < 9; ++i) *reinterpret_cast<T *>(sp + p.ld_off[i]) = f[opp_of(i)];
        } else {
            // natural -> swapped on the rim: periodic wrap by index arithmetic, bounce-back from the own opposite slot
            T *loc[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                if (aa_outside<BC>(i, k, l, p.lnx, p.lny))
                    loc[i] = own + (long long)opp_of(i) * p.pop_stride;
                else
                    loc[i] = buf + (long long)i * p.pop_stride + (long long)(wrap_idx(k - cx_of(i), p.lnx) + 1) * p.pitch +
                             (wrap_idx(l - cy_of(i), p.lny) + PAD_L);
                f[i] = *loc[i];
            }
            aa_lid<T, BC>(p, stash, f, k, l);
            d2q9_collide<T, EXACT>(f, p.omega);
#pragma unroll
            for (int i = 0; i < 9; ++i) *loc[i] = f[opp_of(i)];
        }
    }
}

// Rows [k_lo, k_hi) of the state in NATURAL order into a dense (9, k_hi - k_lo, lny) staging array (downloads
// while the buffer is in the swapped layout).
template <typename T>
__global__ void aa_gather_rows_kernel(const __grid_constant__ StepParams<T> p, T *__restrict__ out, int k_lo, int k_hi)
{
    const long long rows = k_hi - k_lo, n = rows * p.lny;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(t / p.lny), l = (int)(t - (long long)r * p.lny);
#pragma unroll
        for (int i = 0; i < 9; ++i) out[(long long)i * n + t] = aa_natural<T>(p, p.buf[0], i, k_lo + r, l);
    }
}

}  // namespace lbm
