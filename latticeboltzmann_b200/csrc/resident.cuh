// Resident multi-step kernel for L2-resident lattices (BASELINE configs[0] and [1]: the 300 x 200 shear wave, the
// 512 x 512 simple_flows cases).
//
// A lattice of a few hundred thousand cells fits the 126 MB L2 twice over, so its time step is not HBM-bound but
// launch- and latency-bound: one fused launch per step costs ~5 us, three launches per step (simple_flows) ~25 us.
// Here ONE cooperative launch advances the block `nsteps` steps: every CTA keeps its cells, a grid-wide barrier
// (one atomic arrival + one acquire poll per CTA) separates the steps, and everything that used to be a separate
// launch runs inside the loop:
//   * the periodic wrap is index arithmetic on the single self-connected block (no ghost frame, no halo push, no
//     flags) -- the ghosts are refreshed once, by the caller, after the launch;
//   * the shear-wave amplitude probe (shear_wave_opt2.py:99) is an in-kernel epilogue: the cells of the probe
//     column leave uy * uy_k in a double-buffered array and one extra CTA reduces it (same fixed tree as
//     shear_probe_kernel) while the workers already compute the next step;
//   * Couette's collide-before-stream order (PoiseuilleFlow.py:107-111) is a phase shift: with s' = SR(C(s)) the
//     kernel keeps c = C(s) between its passes (c' = C(SR(c))), i.e. one in-place collision first, fused passes,
//     and a last pass without collision -- the stored state between lb_step calls is always the reference's s;
//   * Poiseuille's pressure columns (PoiseuilleFlow.py:76-88: rows 0 and X rewritten from rows X-1 and 1 at the
//     START of a step) are produced at the END of the previous pass by the threads that own rows 0 and X: the plain
//     values of those rows would never be read, so instead their threads recompute the new cells (X-1, l) / (1, l)
//     (one ordinary cell update, the same dependent chain as everybody else) and store the pressure column.  Only the first pass of a launch
//     rewrites the columns in a phase of its own, and the last pass leaves the plain rows, so the stored state
//     between lb_step calls is the reference's.
// Arithmetic, boundary rules and their order are the functions of step_kernel.cuh / temporal.cuh, so results are
// bit-identical to single fused steps.
#pragma once
#include "temporal.cuh"

namespace lbm {

// Threads per CTA of the resident kernel.  Every CTA costs one serialised atomic arrival per barrier, so fewer,
// fatter CTAs shorten the barrier; measured (300 x 200 / 512^2 simple_flows) in profiles/r02_small_lattices.log.
#ifndef LBM_RES_RED_RELEASE
#define LBM_RES_RED_RELEASE 1
#endif
#ifndef LBM_RES_THREADS
#define LBM_RES_THREADS 512
#endif
constexpr int RES_THREADS = LBM_RES_THREADS;

struct ResidentArgs {
    long long nsteps;
    unsigned long long bar_base;    // DevState::grid_bar when the launch starts
    int couette_shift;              // 1: collide in place first, last pass without collision
    int probe_l;                    // local column of the shear probe, -1: none
    const void *uy_k;               // lnx values
    void *series;                   // capacity values
    long long capacity;
    unsigned long long step0;       // step count when the probe was enabled
    void *prod;                     // 2 * lnx values: uy(k, probe_l) * uy_k[k] of the last two steps
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of the (cooperative, hence co-resident) grid: arrive, then poll the counter until `target` arrivals were
// counted.  (Measured alternative: the last arrival publishes a release word in a cache line of its own and the
// others poll that -- one more serial hop, 300 x 200: 3.55 instead of 3.0 us per step -- so the counter is polled.)
__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
#if LBM_RES_RED_RELEASE
        // one fire-and-forget release reduction: the CTA's stores (ordered before it by the bar.sync above, release is
        // cumulative) become visible before the arrival counts, without a separate fence and without the atomic's return trip
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ull) : "memory");
#else
        __threadfence();
        atomicAdd(bar, 1ull);
#endif
        while (ld_acquire_gpu(bar) < target) {}
    }
    __syncthreads();
}

// level n: the current buffer of a single self-connected block, periodic wrap by index arithmetic
template <typename T>
struct SrcWrap {
    const StepParams<T> &p;
    const T *src;
    int k, l;
    template <int I>
    __device__ __forceinline__ T pull() const
    {
        int kk = k - cx_of(I), ll = l - cy_of(I);
        if (cx_of(I) == 1 && kk < 0) kk = p.lnx - 1;
        if (cx_of(I) == -1 && kk >= p.lnx) kk = 0;
        if (cy_of(I) == 1 && ll < 0) ll = p.lny - 1;
        if (cy_of(I) == -1 && ll >= p.lny) ll = 0;
        return __ldcg(src + (long long)I * p.pop_stride + (long long)(kk + 1) * p.pitch + (ll + PAD_L));
    }
    __device__ __forceinline__ T own(int i) const { return __ldcg(src + (long long)i * p.pop_stride + (long long)(k + 1) * p.pitch + (l + PAD_L)); }
};

template <typename T, int BC, bool EXACT>
__global__ void __launch_bounds__(RES_THREADS) resident_kernel(const __grid_constant__ StepParams<T> p, const ResidentArgs a)
{
    __shared__ T red[RES_THREADS];
    DevState *st = p.st;
    const unsigned long long step = *(volatile unsigned long long *)&st->step;
    int par = (int)*(volatile unsigned int *)&st->cur;
    const bool probe = a.probe_l >= 0;
    const int workers = (int)gridDim.x - (probe ? 1 : 0);
    const bool reducer = probe && (int)blockIdx.x == workers;          // the extra CTA: probe reductions only
    const long long n = (long long)p.lnx * p.lny;
    const long long stride = (long long)workers * RES_THREADS;
    const long long t0 = (long long)blockIdx.x * RES_THREADS + threadIdx.x;
    unsigned long long bar = a.bar_base;
    const T *uy_k = static_cast<const T *>(a.uy_k);
    T *prod = static_cast<T *>(a.prod);
    T *series = static_cast<T *>(a.series);

    if (a.couette_shift) {              // c = C(s): in place on the current buffer
        if (!reducer)
            for (long long t = t0; t < n; t += stride) {
                const int k = (int)((unsigned)t / (unsigned)p.lny), l = (int)t - k * p.lny;      // at most 2^20 cells: 32-bit division
                T *q = p.buf[par] + (long long)(k + 1) * p.pitch + (l + PAD_L);
                T f[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) f[i] = __ldcg(q + i * p.pop_stride);
                sf_collide<T>(f, p.omega);
#pragma unroll
                for (int i = 0; i < 9; ++i) q[i * p.pop_stride] = f[i];
            }
        bar += gridDim.x;
        grid_barrier(&st->grid_bar, bar);
    }

    for (long long s = 0; s < a.nsteps; ++s) {
        const T *__restrict__ src = p.buf[par];
        T *__restrict__ dst = p.buf[par ^ 1];
        const bool last = s == a.nsteps - 1;
        if (BC == BC_SF_POISEUILLE && s == 0) {   // pressure columns of the stored state: a phase of its own, once per launch
            if (!reducer)
                for (long long t = t0; t < p.lny; t += stride) sf_pressure_cell<T>(p, p.buf[par], (int)t);
            bar += gridDim.x;
            grid_barrier(&st->grid_bar, bar);
        }
        const bool collide = !(a.couette_shift && last);
        if (!reducer) {
            for (long long t = t0; t < n; t += stride) {
                const int kown = (int)((unsigned)t / (unsigned)p.lny), l = (int)t - kown * p.lny;      // at most 2^20 cells: 32-bit division
                // Poiseuille, all passes but the last: the plain values of rows 0 and X are never read (the pressure columns
                // of the next step replace them), so their threads compute the pressure columns instead -- row 0 from the
                // NEW cell (X-1, l), row X from the new cell (1, l), which they recompute themselves (one ordinary cell
                // update, so every thread of the pass has the same dependent chain)
                const bool pcol = BC == BC_SF_POISEUILLE && !last && (kown == 0 || kown == p.lnx - 1);
                const int k = pcol ? (kown == 0 ? p.lnx - 2 : 1) : kown;
                if (BC == BC_SF_TABLE && p.tab_n > 0 && table_has(p.tab_mask, t)) {               // a cell of the boundary table: its own gather
                    table_cell<T>(p, src, dst, table_index<T>(p, t));
                    continue;
                }
                const SrcWrap<T> sw{p, src, k, l};
                T f[9];
                pull9<T>(sw, f);
                const long long c = (long long)(k + 1) * p.pitch + (l + PAD_L);
                if (BC >= BC_SF_COUETTE)
                    sf_wall_rules<T, BC>(p, src, c, k, l, f);
                else
                    wall_rules<T, BC>(p, sw, f, k, l);
                if (collide) {
                    if (BC >= BC_SF_COUETTE)
                        sf_collide<T>(f, p.omega);
                    else
                        d2q9_collide<T, EXACT>(f, p.omega);
                }
                if (pcol) {
                    // PoiseuilleFlow.py:78-88 for the NEXT step, from the new populations of cell (k, l) (== sf_pressure_cell)
                    T e[9], en[9], rho, ux, uy;
                    sf_moments<T>(f, rho, ux, uy);
                    sf_equilibrium<T>(rho, ux, uy, e);
                    sf_equilibrium<T>(kown == 0 ? p.rho_in : p.rho_out, ux, uy, en);
#pragma unroll
                    for (int i = 0; i < 9; ++i) f[i] = rn_add(en[i], rn_sub(f[i], e[i]));
                }
                T *dp = dst + (long long)(kown + 1) * p.pitch + (l + PAD_L);
#pragma unroll
                for (int i = 0; i < 9; ++i) dp[i * p.pop_stride] = f[i];
                if (probe && l == a.probe_l) {
                    T r, x, y;
                    d2q9_moments<T>(f, r, x, y);
                    prod[(s & 1) * p.lnx + k] = y * uy_k[k];
                }
            }
        }
        bar += gridDim.x;
        grid_barrier(&st->grid_bar, bar);
        if (reducer) {                  // shear_wave_opt2.py:99 for the step just completed (same tree as shear_probe_kernel)
            T acc = T(0);
            for (int k = threadIdx.x; k < p.lnx; k += RES_THREADS) acc += __ldcg(prod + (s & 1) * p.lnx + k);
            red[threadIdx.x] = acc;
            __syncthreads();
            for (int w = RES_THREADS / 2; w > 0; w >>= 1) {
                if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                const long long idx = (long long)(step + (unsigned long long)s + 1ull - a.step0) - 1;
                if (idx >= 0 && idx < a.capacity) series[idx] = red[0] * T(2) / T(p.gnx);
            }
            __syncthreads();
        }
        par ^= 1;
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) {      // after the last barrier: every cell of the final state is stored
        const unsigned long long done = step + (unsigned long long)a.nsteps;
        *(volatile unsigned int *)&st->cur = (unsigned int)par;
        *(volatile unsigned long long *)&st->step = done;
        for (int d = 0; d < NUM_DIRS; ++d) st->flag_in[d] = done;   // the caller refreshes the ghosts next (stream order)
    }
}

// ---- two steps per grid barrier (tiny lattices) ---------------------------------------------------------------
// Measured at 300 x 200 (one cell per thread): the grid barrier alone costs 1.36 us (three dependent L2 round
// trips: arrival, counter update, poll), one step's load -> collide -> store chain 1.58 us; they add up (3.0 us).
// Here a CTA owns a tile of RES2_TX x RES2_TY cells and advances it TWO steps per barrier: level n+1 of the tile
// plus a one-cell halo (wrapped periodically -- the halo cells are real cells of the lattice, computed with their
// own wall rules, which is exactly the reference's roll-then-overwrite semantics) goes to shared memory, level n+2
// of the tile is gathered from there.  The halo is computed redundantly ((TX+2)(TY+2) = 612 cells for 512), so this
// only pays while the barrier dominates: lattices whose tiles all fit the GPU at one CTA per SM.  (Measured: letting a
// CTA walk through several tiles per pass on larger L2-resident lattices does not help -- cavity 512^2 7.09 vs 6.96 us
// per step, 1024^2 28.7 vs 26.1 -- at 20 warps per SM the two dependent phases per tile are latency-bound.)
constexpr int RES2_TX = 16, RES2_TY = 32;
constexpr int RES2_HX = RES2_TX + 2, RES2_HY = RES2_TY + 2, RES2_PITCH = RES2_HY + 2;
constexpr int RES2_THREADS = 640;                   // >= RES2_HX * RES2_HY = 612: one level-(n+1) cell per thread
static_assert(RES2_HX * RES2_HY <= RES2_THREADS && RES2_TX * RES2_TY <= RES2_THREADS, "one cell per thread at both levels");

// level n+1: the tile's shared-memory copy (own cell at [r][c], halo included)
template <typename T>
struct SrcTile {
    const T *lvl1;          // [9][RES2_HX][RES2_PITCH]
    int r, c;
    template <int I>
    __device__ __forceinline__ T pull() const { return lvl1[(I * RES2_HX + (r - cx_of(I))) * RES2_PITCH + (c - cy_of(I))]; }
    __device__ __forceinline__ T own(int i) const { return lvl1[(i * RES2_HX + r) * RES2_PITCH + c]; }
};

template <typename T, int BC, bool EXACT>
__global__ void __launch_bounds__(RES2_THREADS) resident2_kernel(const __grid_constant__ StepParams<T> p, const ResidentArgs a)
{
    __shared__ T lvl1[9 * RES2_HX * RES2_PITCH];
    __shared__ T red[256];
    DevState *st = p.st;
    const unsigned long long step = *(volatile unsigned long long *)&st->step;
    int par = (int)*(volatile unsigned int *)&st->cur;
    const bool probe = a.probe_l >= 0;
    const int workers = (int)gridDim.x - (probe ? 1 : 0);
    const bool reducer = probe && (int)blockIdx.x == workers;
    const int tiles_l = (p.lny + RES2_TY - 1) / RES2_TY;
    const int tk = (int)blockIdx.x / tiles_l, tl = (int)blockIdx.x - tk * tiles_l;
    const int k0 = tk * RES2_TX, l0 = tl * RES2_TY;
    const int txe = min(RES2_TX, p.lnx - k0), tye = min(RES2_TY, p.lny - l0);      // real extent of this tile
    const int t = threadIdx.x;
    unsigned long long bar = a.bar_base;
    const T *uy_k = static_cast<const T *>(a.uy_k);
    T *prod = static_cast<T *>(a.prod);             // [2 passes][2 levels][lnx]
    T *series = static_cast<T *>(a.series);
    const long long passes = a.nsteps / 2;

    for (long long q = 0; q < passes; ++q) {
        const T *__restrict__ src = p.buf[par];
        T *__restrict__ dst = p.buf[par ^ 1];
        if (!reducer) {
            // level n -> n+1 on the tile and its one-cell halo
            const int r = t / RES2_HY, c = t - r * RES2_HY;
            if (r <= txe + 1 && c <= tye + 1 && t < RES2_HX * RES2_HY) {
                const int k = wrap_idx(k0 - 1 + r, p.lnx), l = wrap_idx(l0 - 1 + c, p.lny);
                const SrcWrap<T> sw{p, src, k, l};
                T f[9];
                pull9<T>(sw, f);
                wall_rules<T, BC>(p, sw, f, k, l);
                d2q9_collide<T, EXACT>(f, p.omega);
#pragma unroll
                for (int i = 0; i < 9; ++i) lvl1[(i * RES2_HX + r) * RES2_PITCH + c] = f[i];
                if (probe && l == a.probe_l && r >= 1 && r <= txe && c >= 1 && c <= tye) {
                    T rr, x, y;
                    d2q9_moments<T>(f, rr, x, y);
                    prod[((q & 1) * 2 + 0) * p.lnx + k] = y * uy_k[k];
                }
            }
            __syncthreads();
            // level n+1 -> n+2 on the tile
            const int r2 = t / RES2_TY + 1, c2 = t - (r2 - 1) * RES2_TY + 1;
            if (t < RES2_TX * RES2_TY && r2 <= txe && c2 <= tye) {
                const int k = k0 + r2 - 1, l = l0 + c2 - 1;
                const SrcTile<T> ss{lvl1, r2, c2};
                T g[9];
                pull9<T>(ss, g);
                wall_rules<T, BC>(p, ss, g, k, l);
                d2q9_collide<T, EXACT>(g, p.omega);
                T *dp = dst + (long long)(k + 1) * p.pitch + (l + PAD_L);
#pragma unroll
                for (int i = 0; i < 9; ++i) dp[i * p.pop_stride] = g[i];
                if (probe && l == a.probe_l) {
                    T rr, x, y;
                    d2q9_moments<T>(g, rr, x, y);
                    prod[((q & 1) * 2 + 1) * p.lnx + k] = y * uy_k[k];
                }
            }
        }
        bar += gridDim.x;
        grid_barrier(&st->grid_bar, bar);           // its leading __syncthreads also protects lvl1 for the next pass
        if (reducer) {
            for (int lev = 0; lev < 2; ++lev) {     // shear_wave_opt2.py:99 for both steps of the pass (tree of shear_probe_kernel)
                T acc = T(0);
                if (t < 256) {
                    for (int k = t; k < p.lnx; k += 256) acc += __ldcg(prod + ((q & 1) * 2 + lev) * p.lnx + k);
                    red[t] = acc;
                }
                __syncthreads();
                for (int w = 128; w > 0; w >>= 1) {
                    if (t < w) red[t] += red[t + w];
                    __syncthreads();
                }
                if (t == 0) {
                    const long long idx = (long long)(step + (unsigned long long)(2 * q + lev) + 1ull - a.step0) - 1;
                    if (idx >= 0 && idx < a.capacity) series[idx] = red[0] * T(2) / T(p.gnx);
                }
                __syncthreads();
            }
        }
        par ^= 1;
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long done = step + 2ull * (unsigned long long)passes;
        *(volatile unsigned int *)&st->cur = (unsigned int)par;
        *(volatile unsigned long long *)&st->step = done;
        for (int d = 0; d < NUM_DIRS; ++d) st->flag_in[d] = done;
    }
}

}  // namespace lbm
