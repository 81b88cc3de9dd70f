// The fused D2Q9 time step: pull-stream + boundary handling + BGK collide +
// halo push into the neighbours' ghost layers, ONE kernel launch per step.
//
// Replaces, per step (cavity_opt2.py:272-277):
//   communicate()            cavity_opt2.py:179-210  (4 blocking MPI Sendrecv)
//   stream_and_bounce_back() cavity_opt2.py:109-177  (8 np.roll + ~30 slice stores)
//   D2Q9.collide()           c/d2q9.h:121-131        (serial C++ loop)
//
// Scheme: post[i,k,l] = pre[i,k-cx,l-cy] read from buffer
// `cur`, boundary predicates on GLOBAL coordinates, collide in registers,
// write the other buffer.  Every block always owns a one-cell ghost frame
// (ghost rows inside the arrays, ghost columns in the contiguous `ycol` side
// arrays, lattice.cuh); cells on the block's rim additionally store the 3 (faces) /
// 1 (corners) populations that leave the block straight into the neighbour's
// ghost frame -- local memory for a self-closed periodic ring, a peer-mapped
// NVLink address for another GPU.  Interior cells never test for wrap-around.
//
// Cross-block ordering uses one monotonically increasing flag per direction
// (DevState::flag_in).  Step n's rim CTAs first wait until all 8 neighbours have
// posted `n` (their step n-1 halos are in my ghosts AND they are done reading
// the ghosts I am about to overwrite), and the last rim CTA to finish posts
// `n+1` to all 8 neighbours.  "Rim" CTAs own exactly the block's perimeter cells
// (the only cells that read ghosts, see a wall, or push halos); they take the
// lowest block indices so they are scheduled first and the interior update --
// predicate-free tiles of rows_per_tile x 256 cells -- overlaps the exchange.
#pragma once
#include "d2q9_math.cuh"
#include "lattice.cuh"

namespace lbm {

enum BoundaryKind { BC_PERIODIC = 0, BC_CAVITY = 1, BC_CAVITY_XPERIODIC = 2, BC_SF_COUETTE = 3, BC_SF_POISEUILLE = 4, BC_SF_SLIDING_LID = 5, BC_SF_TABLE = 6 };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Fence at the scope the block's neighbours need: system scope only when a neighbour lives on
// another GPU / in another process; a self-closed ring or blocks sharing the device need gpu scope
// (a system-scope fence costs microseconds, which is the whole budget of a small lattice's step).
__device__ __forceinline__ void halo_fence(int sys_scope)
{
    if (sys_scope)
        __threadfence_system();
    else
        __threadfence();
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- state access that is aware of the in-place (AA) layout, aa.cuh -------------------------------------
// A lattice created with LB_CREATE_INPLACE keeps ONE buffer whose layout alternates between natural (f*_i(x) in
// slot i of cell x) and swapped (f*_i(x) in slot opp(i) of cell x + c_i, or in its natural place where x + c_i
// lies outside a wall).  Read-only consumers (moments, digest, downloads) go through aa_natural().
__device__ __forceinline__ bool aa_outside_rt(int bc, int i, int k, int l, int lnx, int lny)
{
    if (bc == BC_PERIODIC) return false;
    const bool walls = bc == BC_CAVITY;
    return (cy_of(i) == 1 && l == 0) || (cy_of(i) == -1 && l == lny - 1) || (walls && cx_of(i) == 1 && k == 0) ||
           (walls && cx_of(i) == -1 && k == lnx - 1);
}
__device__ __forceinline__ int wrap_idx(int a, int n) { return a < 0 ? a + n : (a >= n ? a - n : a); }

// pre[i](k, l) = f*_i at cell (k, l), whichever layout the single buffer is in (see the file header)
template <typename T>
__device__ __forceinline__ T aa_natural(const StepParams<T> &p, const T *buf, int i, int k, int l)
{
    int kk = k, ll = l, slot = i;
    if (p.aa_swapped && i != 0 && !aa_outside_rt(p.bc, opp_of(i), k, l, p.lnx, p.lny)) {
        kk = wrap_idx(k + cx_of(i), p.lnx);
        ll = wrap_idx(l + cy_of(i), p.lny);
        slot = opp_of(i);
    }
    return __ldcg(buf + (long long)slot * p.pop_stride + (long long)(kk + 1) * p.pitch + (ll + PAD_L));
}

// Loads: rim tiles read ghost cells that a peer may have written during this
// launch's lifetime, so they bypass the (non-coherent) L1.
template <bool RIM, typename T>
__device__ __forceinline__ T ld_f(const T *p)
{
    if (RIM) return __ldcg(p);
    return *p;
}

// Perimeter cell t of the block -> (k, l).  Segments: row k=0, row k=lnx-1,
// column l=0 (k in 1..lnx-2), column l=lny-1 (k in 1..lnx-2).
template <typename T>
__device__ __forceinline__ void decode_perimeter(const StepParams<T> &p, long long t, int &k, int &l)
{
    if (p.all_rim) { k = (int)(t / p.lny); l = (int)(t - (long long)k * p.lny); return; }
    const long long a = p.lny, b = (p.lnx > 1) ? p.lny : 0, c = (p.lnx > 2) ? p.lnx - 2 : 0;
    if (t < a) { k = 0; l = (int)t; }
    else if (t < a + b) { k = p.lnx - 1; l = (int)(t - a); }
    else if (t < a + b + c) { k = 1 + (int)(t - a - b); l = 0; }
    else { k = 1 + (int)(t - a - b - c); l = p.lny - 1; }
}

// Store the populations that leave the block through cell (k, l) into the
// neighbours' ghost frames of buffer `par`.
template <typename T>
__device__ __forceinline__ void push_halo(const StepParams<T> &p, int par, int k, int l, const T (&f)[9])
{
    const bool xl = (k == 0), xh = (k == p.lnx - 1), yl = (l == 0), yh = (l == p.lny - 1);
    if (!(xl | xh | yl | yh)) return;
#define LBM_PUSH(D, KK, LL, I)                                                              \
    do {                                                                                    \
        const NbrView<T> &nb = p.nbr[D];                                                    \
        nb.buf[par][(long long)(I) * nb.pop_stride + (long long)((KK) + 1) * nb.pitch + ((LL) + PAD_L)] = f[I]; \
    } while (0)
    if (xh) {   // +x face -> ghost row -1 of the right neighbour: E, NE, SE
        LBM_PUSH(1, -1, l, QE);
        LBM_PUSH(1, -1, l, QNE);
        LBM_PUSH(1, -1, l, QSE);
    }
    if (xl) {   // -x face -> ghost row lnx of the left neighbour: W, NW, SW
        LBM_PUSH(0, p.nbr[0].lnx, l, QW);
        LBM_PUSH(0, p.nbr[0].lnx, l, QNW);
        LBM_PUSH(0, p.nbr[0].lnx, l, QSW);
    }
#define LBM_PUSH_Y(D, SIDE, KK, I)                                                          \
    do {                                                                                    \
        const NbrView<T> &nb = p.nbr[D];                                                    \
        nb.ycol[par][((SIDE) * 3 + ycol_slot(I)) * (long long)(nb.lnx + 2) + ((KK) + 1)] = f[I]; \
    } while (0)
    if (yh) {   // +y face -> ghost column -1 (ycol side 0) of the upper neighbour: N, NE, NW
        LBM_PUSH_Y(3, 0, k, QN);
        LBM_PUSH_Y(3, 0, k, QNE);
        LBM_PUSH_Y(3, 0, k, QNW);
    }
    if (yl) {   // -y face -> ghost column lny (ycol side 1) of the lower neighbour: S, SW, SE
        LBM_PUSH_Y(2, 1, k, QS);
        LBM_PUSH_Y(2, 1, k, QSW);
        LBM_PUSH_Y(2, 1, k, QSE);
    }
    if (xh && yh) LBM_PUSH_Y(7, 0, -1, QNE);
    if (xh && yl) LBM_PUSH_Y(6, 1, -1, QSE);
    if (xl && yh) LBM_PUSH_Y(5, 0, p.nbr[5].lnx, QNW);
    if (xl && yl) LBM_PUSH_Y(4, 1, p.nbr[4].lnx, QSW);
#undef LBM_PUSH_Y
#undef LBM_PUSH
}

// simple_flows flavour: explicit wall LAYERS at l = 0 / T (and k = 0 / X for the sliding
// lid); after the periodic pull, populations that entered a wall layer are copied back into the adjacent fluid
// row/column -- written here as a gather from the pre-stream state `src` (c = the cell's own element offset).
// Single block only, so local == global coordinates.
template <typename T, int BC>
__device__ __forceinline__ void sf_wall_rules(const StepParams<T> &p, const T *__restrict__ src, long long c, int k, int l, T (&f)[9])
{
    if (BC == BC_SF_TABLE) return;      // walls come from the per-cell table (table_cell), everything else streams periodically
    const long long S = p.pop_stride, P = p.pitch;
        const int X = p.lnx - 1, Tt = p.lny - 1;
#define LBM_PRE(I, DK, DL) ld_f<true>(src + (long long)(I) * S + c + (long long)(DK) * P + (DL))
        if (BC == BC_SF_SLIDING_LID && l >= 1 && l <= Tt - 1) {     // slidingLid.py:72-78
            if (k == 1) { f[QE] = LBM_PRE(QW, 0, 0); f[QNE] = LBM_PRE(QSW, 0, 1); f[QSE] = LBM_PRE(QNW, 0, -1); }
            if (k == X - 1) { f[QW] = LBM_PRE(QE, 0, 0); f[QNW] = LBM_PRE(QSE, 0, 1); f[QSW] = LBM_PRE(QNE, 0, -1); }
        }
        if (k >= 1 && k <= X - 1) {                                  // PoiseuilleFlow.py:65-74, slidingLid.py:83-91
            if (l == 1) { f[QN] = LBM_PRE(QS, 0, 0); f[QNE] = LBM_PRE(QSW, 1, 0); f[QNW] = LBM_PRE(QSE, -1, 0); }
            if (l == Tt - 1) {
                const T g2 = LBM_PRE(QN, 0, 0), g5 = LBM_PRE(QNE, -1, 0), g6 = LBM_PRE(QNW, 1, 0);
                T shift = p.sf_uw6;
                if (BC == BC_SF_SLIDING_LID) {                      // rho_wall, slidingLid.py:87-91
                    const T g0 = LBM_PRE(Q0, 0, 1), g1 = LBM_PRE(QE, -1, 1), g3 = LBM_PRE(QW, 1, 1);
                    const T rw = rn_add(rn_add(rn_add(rn_mul(T(2), rn_add(rn_add(g2, g5), g6)), g0), g1), g3);
                    shift = rn_mul(p.sf_uw6, rw);
                }
                f[QS] = g2;
                f[QSW] = rn_sub(g5, shift);
                f[QSE] = rn_add(g6, shift);
            }
        }
#undef LBM_PRE
}

// One cell: gather, boundary rules, collide, store (+ halo push on rim tiles).
template <typename T, int BC, bool EXACT, bool COLLIDE, bool RIM>
__device__ __forceinline__ void update_cell(const StepParams<T> &p, const T *__restrict__ src, T *__restrict__ dst,
                                            int par_dst, int k, int l)
{
    const long long S = p.pop_stride, P = p.pitch;
    const long long c = (long long)(k + 1) * P + (l + PAD_L);
    const char *sp = reinterpret_cast<const char *>(src + c);
    T f[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        if (RIM && cy_of(i) == 1 && l == 0)                  // source (k-cx, -1): ghost column below
            f[i] = __ldcg(p.ycol[par_dst ^ 1] + (0 * 3 + ycol_slot(i)) * (long long)(p.lnx + 2) + (k - cx_of(i) + 1));
        else if (RIM && cy_of(i) == -1 && l == p.lny - 1)    // source (k-cx, lny): ghost column above
            f[i] = __ldcg(p.ycol[par_dst ^ 1] + (1 * 3 + ycol_slot(i)) * (long long)(p.lnx + 2) + (k - cx_of(i) + 1));
        else
            f[i] = ld_f<RIM>(reinterpret_cast<const T *>(sp + p.ld_off[i]));
    }

    if (RIM && BC >= BC_SF_COUETTE) {
        sf_wall_rules<T, BC>(p, src, c, k, l, f);
    } else if (RIM && BC != BC_PERIODIC) {
        // SURVEY.md App. A.2 == cavity_opt2.py:133-177 as a gather.  Predicates on
        // global coordinates; the wall sits half a cell outside the lattice.
        const long long gk = p.x0 + k, gl = p.y0 + l;
        const bool walls = (BC == BC_CAVITY);
        const bool bottom = (gl == 0), top = (gl == p.gny - 1);
        const bool left = walls && (gk == 0), right = walls && (gk == p.gnx - 1);
        if (bottom | top | left | right) {
            T lid = T(0), o_nw = T(0), o_ne = T(0);
            if (top) {
                // cavity_opt2.py:140-142: rho from 3 pre-stream + 6 streamed (periodically
                // rolled, pre-wall-overwrite) populations, summed left to right.
                o_nw = ld_f<true>(src + QNW * S + c);
                o_ne = ld_f<true>(src + QNE * S + c);
                T rho = rn_add(o_nw, ld_f<true>(src + QN * S + c));
                rho = rn_add(rho, o_ne);
                rho = rn_add(rho, f[QNW]);
                rho = rn_add(rho, f[QN]);
                rho = rn_add(rho, f[QNE]);
                rho = rn_add(rho, f[QW]);
                rho = rn_add(rho, f[Q0]);
                rho = rn_add(rho, f[QE]);
                // 6*w_i[D.SE]*rho*u0 (:144-145): w = fp(1/36) rounded to T, 6*w rounded in T.
                const T six_w = rn_mul(T(6), T(1.0 / 36.0));
                lid = rn_mul(rn_mul(six_w, rho), p.u_wall);
            }
            // half-way bounce-back: a population whose source cell lies outside the box is
            // replaced by the cell's own opposite pre-stream population (:134-136,:150-177).
#pragma unroll
            for (int i = 1; i < 9; ++i) {
                const bool outside = (cy_of(i) == 1 && bottom) || (cy_of(i) == -1 && top) ||
                                     (cx_of(i) == 1 && left) || (cx_of(i) == -1 && right);
                if (outside) f[i] = ld_f<true>(src + opp_of(i) * S + c);
            }
            if (top) {
                if (!left) f[QSE] = rn_add(o_nw, lid);    // :144, overridden by the left wall at k=0 (:152,:172)
                if (!right) f[QSW] = rn_sub(o_ne, lid);   // :145, overridden by the right wall at k=X (:157,:177)
            }
        }
    }

    if (COLLIDE) {
        if (BC >= BC_SF_COUETTE)
            sf_collide<T>(f, p.omega);
        else
            d2q9_collide<T, EXACT>(f, p.omega);
    }

    // stores walk the populations with one stride (a 9-entry offset table would be hoisted into 18 registers)
    T *dp = dst + c;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        *dp = f[i];
        dp += S;
    }
    if (RIM) push_halo<T>(p, par_dst, k, l, f);
}

// Resident CTAs per SM the register allocation is tuned for (x 256 threads).
// Resident CTAs per SM the register allocation targets (x 256 threads).  With the 3-operation
// constant divisions the EXACT fp64 kernel fits 62 registers without spills, so every variant
// runs 4 CTAs (32 warps) per SM.  (Measured alternatives at 4096^2 fp64 EXACT: 3 CTAs 44.0-44.3,
// 4 CTAs 44.0-45.0, 2 CTAs 36-38, software-prefetched rows 44.5-45.2 GLUPS -- all within noise of
// the HBM roofline except 2 CTAs, so the simplest loop is kept.)
#ifndef LBM_EXACT_MINB
#define LBM_EXACT_MINB 4
#endif
#ifndef LBM_FAST_MINB
#define LBM_FAST_MINB 4
#endif
template <typename T, bool EXACT>
__host__ __device__ constexpr int min_ctas_per_sm() { return (sizeof(T) == 8 && EXACT) ? LBM_EXACT_MINB : LBM_FAST_MINB; }

// Interior cell, split into its load and its compute+store half so that the row loop can be
// software-pipelined (loads of row k+1 in flight while row k is collided).
#ifndef LBM_LD_HINT
#define LBM_LD_HINT 0      // 0: default caching, 1: ld.global.cs (streaming), 2: ld.global.cg (L2 only)
#endif
#ifndef LBM_ST_HINT
#define LBM_ST_HINT 0      // 0: default (write-back), 1: st.global.cs (streaming), 2: st.global.cg
#endif
template <typename T>
__device__ __forceinline__ void interior_load(const StepParams<T> &p, const char *sp, T (&f)[9])
{
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const T *q = reinterpret_cast<const T *>(sp + p.ld_off[i]);
        f[i] = LBM_LD_HINT == 1 ? __ldcs(q) : (LBM_LD_HINT == 2 ? __ldcg(q) : *q);
    }
}
template <typename T, int BC, bool EXACT, bool COLLIDE>
__device__ __forceinline__ void interior_finish(const StepParams<T> &p, T *dp, T (&f)[9])
{
    if (COLLIDE) {
        if (BC >= BC_SF_COUETTE)
            sf_collide<T>(f, p.omega);
        else
            d2q9_collide<T, EXACT>(f, p.omega);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        if (LBM_ST_HINT == 1)
            __stcs(dp, f[i]);
        else if (LBM_ST_HINT == 2)
            __stcg(dp, f[i]);
        else
            *dp = f[i];
        dp += p.pop_stride;
    }
}

// ---- rim CTAs: the block's perimeter cells, ghost reads, halo pushes ----
template <typename T, int BC, bool EXACT, bool COLLIDE>
__device__ __forceinline__ void rim_cta(const StepParams<T> &p, DevState *st, unsigned long long step, const T *__restrict__ src,
                                        T *__restrict__ dst, int par)
{
    // Wait for the 8 neighbours' step-(n-1) halos (and for them to be done reading the ghosts this
    // step overwrites).
    if (threadIdx.x < NUM_DIRS) {
        if (ld_acquire_sys(&st->flag_in[threadIdx.x]) < step && *(volatile unsigned int *)&st->error == 0) {
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(&st->flag_in[threadIdx.x]) < step) {
                if (global_timer_ns() - t0 > p.halo_timeout_ns) {
                    atomicExch(&st->error, 1u);
                    break;
                }
            }
        }
    }
    __syncthreads();
    const long long t = (long long)blockIdx.x * TILE_L + threadIdx.x;
    if (t < p.n_perimeter) {
        int k, l;
        decode_perimeter<T>(p, t, k, l);
        update_cell<T, BC, EXACT, COLLIDE, true>(p, src, dst, par ^ 1, k, l);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        halo_fence(p.sys_scope);                  // this CTA's pushes are visible before it is counted
        const unsigned int prev = atomicAdd(&st->edge_done, 1u);
        if (prev == (unsigned int)p.n_rim_ctas - 1u) {
            st->edge_done = 0u;
            halo_fence(p.sys_scope);              // fence + relaxed stores = release of all rim CTAs' pushes
#pragma unroll
            for (int d = 0; d < NUM_DIRS; ++d) st_relaxed_sys(p.nbr[d].flag_in + dir_opp(d), step + 1ull);
        }
    }
}

// The last CTA of the launch publishes the new step count (read by the next launch -- which makes
// the launch arguments step-independent and the whole loop CUDA-graph replayable).
__device__ __forceinline__ void publish_step(DevState *st, unsigned long long step, int par)
{
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&st->all_done, 1u);
        if (prev == gridDim.x - 1u) {
            st->all_done = 0u;
            *(volatile unsigned int *)&st->cur = (unsigned int)(par ^ 1);
            *(volatile unsigned long long *)&st->step = step + 1ull;
        }
    }
}

template <typename T, int BC, bool EXACT, bool COLLIDE>
__global__ void __launch_bounds__(TILE_L, min_ctas_per_sm<T, EXACT>()) step_kernel(const __grid_constant__ StepParams<T> p)
{
    DevState *st = p.st;
    const unsigned long long step = *(volatile unsigned long long *)&st->step;
    const int par = (int)*(volatile unsigned int *)&st->cur;
    const T *__restrict__ src = p.buf[par];
    T *__restrict__ dst = p.buf[par ^ 1];

    if ((int)blockIdx.x < p.n_rim_ctas) {
        rim_cta<T, BC, EXACT, COLLIDE>(p, st, step, src, dst, par);
    } else {
        // ---- interior CTAs: rows_per_tile x 256 cells, no ghosts, no predicates ----
        const int tile = (int)blockIdx.x - p.n_rim_ctas;
        const int kt = tile / p.tiles_l, lt = tile - kt * p.tiles_l;
        const int l = lt * TILE_L + threadIdx.x;
        const int k0 = max(kt * p.rows_per_tile, 1);
        const int k1 = min(kt * p.rows_per_tile + p.rows_per_tile, p.lnx - 1);
        if (l >= 1 && l < p.lny - 1 && k0 < k1) {
            const long long c0 = (long long)(k0 + 1) * p.pitch + (l + PAD_L);
            const char *sp = reinterpret_cast<const char *>(src + c0);
            T *dp = dst + c0;
            const long long row_bytes = p.pitch * (long long)sizeof(T);
            int k = k0;
            if (sizeof(T) == 4) {
                // fp32 moves half the bytes per lane, so one row per iteration leaves too few bytes in
                // flight per warp (measured 77.8 GLUPS); two rows per iteration = 18 loads in flight per
                // thread, the same 72 B as fp64 (85.8 GLUPS).  For fp64 the single-row loop is as fast or faster.
#pragma unroll 1
                for (; k + 1 < k1; k += 2) {
                    T fa[9], fb[9];
                    interior_load<T>(p, sp, fa);
                    interior_load<T>(p, sp + row_bytes, fb);
                    interior_finish<T, BC, EXACT, COLLIDE>(p, dp, fa);
                    interior_finish<T, BC, EXACT, COLLIDE>(p, dp + p.pitch, fb);
                    sp += 2 * row_bytes;
                    dp += 2 * p.pitch;
                }
            }
#pragma unroll 1
            for (; k < k1; ++k) {
                T f[9];
                interior_load<T>(p, sp, f);
                interior_finish<T, BC, EXACT, COLLIDE>(p, dp, f);
                sp += row_bytes;
                dp += p.pitch;
            }
        }
        __syncthreads();
    }
    publish_step(st, step, par);
}

// Rows [k_lo, k_hi) of one step through the general (rim) cell path, without the flag protocol and
// without touching the step counter: the building block of the slab-pipelined host step
// (lb_step_host), where the host orders H2D -> compute -> D2H per slab with events.
template <typename T, int BC, bool EXACT>
__global__ void __launch_bounds__(TILE_L) step_rows_kernel(const __grid_constant__ StepParams<T> p, int k_lo, int k_hi)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    const T *__restrict__ src = p.buf[par];
    T *__restrict__ dst = p.buf[par ^ 1];
    const int tiles_l = (p.lny + TILE_L - 1) / TILE_L;
    const int k = k_lo + (int)blockIdx.x / tiles_l;
    const int l = ((int)blockIdx.x % tiles_l) * TILE_L + threadIdx.x;
    if (k < k_hi && l < p.lny) update_cell<T, BC, EXACT, true, true>(p, src, dst, par ^ 1, k, l);
}

// Completes a host-ordered step of a self-connected block: bump the step counter and post the
// block's own halo flags so that a following fused (flag-ordered) step finds them satisfied.
__global__ void advance_step_kernel(DevState *st)
{
    const unsigned long long s = st->step + 1ull;
    st->step = s;
    st->cur ^= 1u;
    for (int d = 0; d < NUM_DIRS; ++d) st->flag_in[d] = s;
}

// Push the rim of the CURRENT buffer (used once after init / upload).
template <typename T>
__global__ void halo_refresh_kernel(const __grid_constant__ StepParams<T> p, int k_lo, int k_hi)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    const T *src = p.buf[par];
    const long long n_rim = 2ll * p.lnx + 2ll * p.lny;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_rim; t += (long long)gridDim.x * blockDim.x) {
        int k, l;
        if (t < p.lny) { k = 0; l = (int)t; }
        else if (t < 2ll * p.lny) { k = p.lnx - 1; l = (int)(t - p.lny); }
        else if (t < 2ll * p.lny + p.lnx) { k = (int)(t - 2ll * p.lny); l = 0; }
        else { k = (int)(t - 2ll * p.lny - p.lnx); l = p.lny - 1; }
        if (k < k_lo || k >= k_hi) continue;      // row-range variant used by the pipelined host step
        T f[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = __ldcg(src + i * p.pop_stride + (long long)(k + 1) * p.pitch + (l + PAD_L));
        push_halo<T>(p, par, k, l, f);
    }
}

// f = feq(rho, ux, uy) on the real cells of the current buffer (c/d2q9.h:98-108).
template <typename T>
__global__ void init_equilibrium_kernel(const __grid_constant__ StepParams<T> p, const T *__restrict__ rho,
                                        const T *__restrict__ ux, const T *__restrict__ uy)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    T *dst = p.buf[par];
    const long long n = (long long)p.lnx * p.lny;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / p.lny), l = (int)(t % p.lny);
        T e[9];
        d2q9_equilibrium<T, true>(rho ? rho[t] : T(1), ux ? ux[t] : T(0), uy ? uy[t] : T(0), e);
#pragma unroll
        for (int i = 0; i < 9; ++i) dst[i * p.pop_stride + (long long)(k + 1) * p.pitch + (l + PAD_L)] = e[i];
    }
}

// rho, ux, uy of the current buffer into dense (lnx, lny) arrays (cavity_opt2.py:280-281).
template <typename T, bool SF>
__global__ void moments_kernel(const __grid_constant__ StepParams<T> p, T *__restrict__ rho, T *__restrict__ ux,
                               T *__restrict__ uy)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    const T *src = p.buf[par];
    const long long n = (long long)p.lnx * p.lny;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / p.lny), l = (int)(t % p.lny);
        T f[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = aa_natural<T>(p, src, i, k, l);
        T r, x, y;
        if (SF)
            sf_moments<T>(f, r, x, y);      // PoiseuilleFlow.py:49-53
        else
            d2q9_moments<T>(f, r, x, y);
        if (rho) rho[t] = r;
        if (ux) ux[t] = x;
        if (uy) uy[t] = y;
    }
}

// BC_SF_TABLE: cell j of the boundary table -- gather the nine populations from their tabulated pre-stream sources
// (+ constant), collide, store (lattice.cuh: tab_*; latticeboltzmann_b200/boundary_table.py builds the table).
// The table itself is read-only for the lifetime of a launch: __ldg keeps it in L1, so the index -> source -> value
// chain costs one L2 round trip (the value), not three.
template <typename T>
__device__ __forceinline__ void table_cell(const StepParams<T> &p, const T *__restrict__ src, T *__restrict__ dst, int j)
{
    const int cell = __ldg(p.tab_cells + j);
    const int k = cell / p.lny, l = cell - k * p.lny;
    T f[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const T v = __ldcg(src + __ldg(p.tab_src + 9 * (long long)j + i));
        const T c = __ldg(p.tab_add + 9 * (long long)j + i);
        f[i] = c == T(0) ? v : rn_add(v, c);      // x + 0 would turn -0 into +0; a plain copy must stay a copy
    }
    sf_collide<T>(f, p.omega);
    T *dp = dst + (long long)(k + 1) * p.pitch + (l + PAD_L);
#pragma unroll
    for (int i = 0; i < 9; ++i) dp[i * p.pop_stride] = f[i];
}
__device__ __forceinline__ bool table_has(const unsigned int *mask, long long cell) { return (__ldg(mask + (cell >> 5)) >> (cell & 31)) & 1u; }
// index of a listed cell in the (cell-sorted) table
template <typename T>
__device__ __forceinline__ int table_index(const StepParams<T> &p, long long cell)
{
    return __ldg(p.tab_rank + (cell >> 5)) + __popc(__ldg(p.tab_mask + (cell >> 5)) & ((1u << (cell & 31)) - 1u));
}

// Per-step path of BC_SF_TABLE: after step_kernel streamed + collided EVERY cell periodically into the new current
// buffer, the listed cells are recomputed from the previous buffer (still intact: A/B) with their table.
template <typename T>
__global__ void table_cells_kernel(const __grid_constant__ StepParams<T> p)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;      // already flipped by step_kernel
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p.tab_n; j += gridDim.x * blockDim.x)
        table_cell<T>(p, p.buf[par ^ 1], p.buf[par], j);
}

// Order-independent 64-bit digest of the current state: sum over populations and real cells of
// mix(bits(f) , GLOBAL index), modulo 2^64.  Because the index is global, the digests of the blocks of any
// decomposition add up to the digest of the undecomposed lattice: equal digests <=> bit-identical fields
// (up to 2^-64 collisions) without moving the field off the device.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
template <typename T>
__global__ void checksum_kernel(const __grid_constant__ StepParams<T> p, unsigned long long *out)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    const T *src = p.buf[par];
    const long long n = (long long)p.lnx * p.lny;
    unsigned long long acc = 0ull;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / p.lny), l = (int)(t % p.lny);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const T v = aa_natural<T>(p, src, i, k, l);
            unsigned long long bits;
            if (sizeof(T) == 8)
                bits = (unsigned long long)__double_as_longlong((double)v);
            else
                bits = (unsigned long long)(unsigned int)__float_as_int((float)v);
            const unsigned long long idx = ((unsigned long long)i * (unsigned long long)p.gnx + (unsigned long long)(p.x0 + k)) * (unsigned long long)p.gny + (unsigned long long)(p.y0 + l);
            acc += mix64(bits ^ mix64(idx + 0x9e3779b97f4a7c15ull));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// simple_flows: in-place collision of the whole current buffer (Couette's step order is
// moments -> collide -> stream -> reflect, PoiseuilleFlow.py:107-111).
template <typename T>
__global__ void sf_collide_inplace_kernel(const __grid_constant__ StepParams<T> p)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    T *buf = p.buf[par];
    const long long n = (long long)p.lnx * p.lny;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / p.lny), l = (int)(t % p.lny);
        T *q = buf + (long long)(k + 1) * p.pitch + (l + PAD_L);
        T f[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = q[i * p.pop_stride];
        sf_collide<T>(f, p.omega);
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i * p.pop_stride] = f[i];
    }
}

// simple_flows: pressure-periodic inlet/outlet (PoiseuilleFlow.py:76-88), in place on rows k = 0 and
// k = X of the current buffer:  f[:,0,l] = feq(rho_in, u[X-1,l]) + (f[:,X-1,l] - feq[:,X-1,l]),
//                               f[:,X,l] = feq(rho_out, u[1,l]) + (f[:,1,l]   - feq[:,1,l]).
template <typename T>
__device__ __forceinline__ void sf_pressure_cell(const StepParams<T> &p, T *buf, int l)
{
    const int X = p.lnx - 1;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int ks = side == 0 ? X - 1 : 1, kd = side == 0 ? 0 : X;
        const T *q = buf + (long long)(ks + 1) * p.pitch + (l + PAD_L);
        T f[9], e[9], en[9], rho, ux, uy;
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = __ldcg(q + i * p.pop_stride);
        sf_moments<T>(f, rho, ux, uy);
        sf_equilibrium<T>(rho, ux, uy, e);
        sf_equilibrium<T>(side == 0 ? p.rho_in : p.rho_out, ux, uy, en);
        T *d = buf + (long long)(kd + 1) * p.pitch + (l + PAD_L);
#pragma unroll
        for (int i = 0; i < 9; ++i) d[i * p.pop_stride] = rn_add(en[i], rn_sub(f[i], e[i]));
    }
}
template <typename T>
__global__ void sf_pressure_kernel(const __grid_constant__ StepParams<T> p)
{
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    T *buf = p.buf[par];
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < p.lny; l += gridDim.x * blockDim.x) sf_pressure_cell<T>(p, buf, l);
}

// shear_wave_opt2.py:99 -- one block; deterministic tree reduction.
template <typename T>
__global__ void shear_probe_kernel(const __grid_constant__ StepParams<T> p, int l_local, const T *__restrict__ uy_k,
                                   T *__restrict__ series, long long capacity, unsigned long long step0)
{
    __shared__ T red[256];
    const unsigned long long step = *(volatile unsigned long long *)&p.st->step;
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    const T *src = p.buf[par];
    T acc = T(0);
    for (int k = threadIdx.x; k < p.lnx; k += blockDim.x) {
        T f[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = src[i * p.pop_stride + (long long)(k + 1) * p.pitch + (l_local + PAD_L)];
        T r, x, y;
        d2q9_moments<T>(f, r, x, y);
        acc += y * uy_k[k];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long idx = (long long)(step - step0) - 1;
        if (idx >= 0 && idx < capacity) series[idx] = red[0] * T(2) / T(p.gnx);
    }
}

}  // namespace lbm
