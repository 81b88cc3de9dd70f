// Temporal blocking: TWO time steps per pass over HBM ("double step").
//
// The single-step kernel is at the DRAM pins (144 B per cell per step, step_kernel.cuh).  The only way
// past that roofline is to touch HBM less: advance a tile two levels while it sits on chip, so that a
// cell costs 72 B read + 72 B written per TWO steps.  Results are bit-identical to two single steps
// (same per-level arithmetic, same boundary rules, same halo values).
//
// One double step n -> n+2 of a block is three launches on the block's stream:
//
//   K1  t2_frame1   level n -> n+1 on the FRAME (cells closer than 3 to the block perimeter), from the
//                   main buffer and its level-n ghosts, into the small frame storage (lattice.cuh);
//                   perimeter cells push their leaving populations into the neighbours' level-(n+1)
//                   frame ghosts (peer stores) and the last CTA posts fflag = n+1.
//   K2  t2_interior level n -> n+2 on the DEEP INTERIOR (distance >= 2): per tile, rows of level n+1
//                   (one-cell halo, recomputed redundantly at tile edges: 2 rows per t2_rows, 2 columns
//                   per 254) go through a 4-row shared-memory ring and are consumed by the level-(n+2)
//                   pull.  No ghosts, no walls, no flags: all sources are real cells of this block.
//   K3  t2_frame2   level n+1 -> n+2 on the cells closer than 2 to the perimeter, from the frame storage
//                   and the level-(n+1) frame ghosts; stores into the main destination buffer, pushes the
//                   level-(n+2) halos into the neighbours' main ghosts (push_halo), posts flag = n+2 and
//                   publishes step = n+2.
//
// Flag protocol (monotone counters in DevState, same reasoning as step_kernel.cuh): K1 waits for
// flag_in >= n (level-n main ghosts present, neighbours done with the frame ghosts K1 overwrites), K3
// waits for fflag_in >= n+1 (level-(n+1) frame ghosts present).
//
// K2 touches neither ghosts nor frame storage nor flags, and K1 / K3 touch no deep-interior cell of the destination
// buffer, so lb_step runs K1 -> K3 on a second (high-priority) stream CONCURRENTLY with K2 and joins the two at the
// end of the pass (lattice_api.cu: launch_passes): the small latency-bound frame kernels and the halo exchange hide
// behind the interior update instead of adding to it.  Whichever of K2 / K3 finishes second publishes the pass
// (t2_part_done): buffer flip + step counter, which the kernels of the NEXT pass read.  Drivers that interleave
// several blocks on one stream (lb_double_step_phase) run the three kernels in order; the same rule then makes K3
// the publisher.
#pragma once
#include "step_kernel.cuh"

namespace lbm {

#ifndef LBM_T2_TILE
#define LBM_T2_TILE 256
#endif
constexpr int T2_TILE = LBM_T2_TILE;  // threads (= level-(n+1) columns incl. the one-cell halo) per fused tile
// Column mapping of a fused tile: tile lt emits the level-(n+2) columns [max(2, T2_W*lt + T2_S), T2_W*(lt+1) + T2_S)
// and thread t works on column T2_W*lt + T2_S - T2_OFF + t (level n+1 on threads T2_OFF-1 .. T2_OFF+T2_W).
//   (254, 2, 1): every thread busy, but tile seams fall 16 B into a 32-byte sector on every other tile;
//   (252, 0, 2): every tile's stores start on a sector boundary (252 * 8 B = 63 sectors), 2 idle lanes.
#ifndef LBM_T2_W
#define LBM_T2_W 254
#endif
#ifndef LBM_T2_S
#define LBM_T2_S 2
#endif
#ifndef LBM_T2_OFF
#define LBM_T2_OFF 1
#endif
constexpr int T2_W = LBM_T2_W, T2_S = LBM_T2_S, T2_OFF = LBM_T2_OFF;
static_assert(T2_OFF >= 1 && T2_OFF + T2_W <= T2_TILE - 1, "level-(n+1) halo columns must fit the tile");
static_assert(T2_S == 0 || T2_S == 2, "tiles start at column 2 (the frame) or on multiples of T2_W");
__host__ __device__ inline int t2_tiles_over(long long lny) { return lny > 4 ? (int)((lny - 2 - T2_S + T2_W - 1) / T2_W) : 0; }
// Measured at 16384^2 fp64 EXACT (GLUPS), round 1: all 9 populations through the ring, 3 CTAs/SM: 61.5; six-population
// ring + register-kept rest/E/W with register prefetch of the next row at 80 registers / 3 CTAs: 58.8; the same
// WITHOUT prefetch at 64 registers / 4 CTAs (32 warps): 78.1; 128-thread tiles: 60-61; 3-slot ring, two barriers per
// row: 49-63.  Round 2 (tools/t2_variants.py, profiles/r02_t2_variants.log): LBM_T2_COMPACT_RING keeps only the rows
// each ring population still needs (18 instead of 24 population-rows): 78.6 -> 80.3 (32-row tiles) / 80.8 (64-row
// tiles); staging the NEXT row's nine level-n sources with per-thread 8-byte cp.async (LDGSTS): 60.6 -- the LSU/MIO
// path of 9 x 8-byte LDGSTS per cell costs more than the latency it hides (ncu: mio_throttle 5.6 per issue), so that
// variant was removed again; sector-aligned tile seams (W=252): 78.4, no gain.
// Launch order / walking direction (profiles/r02_t2_order_sweep.log): vertically adjacent tiles walking towards their
// common seam (odd bands top-down) so that the shared halo rows are read at the same time and hit L2, with bands
// launched in groups of 1 / 2 / 4 / 8 / 16 / 32 / all: 86.7 / 83.6 / 85.8 / 85.6 / 83.4 / 84.1 / 75.6 vs 86.3 GLUPS for the
// plain band-major upward walk at 16384^2 (4096^2: 76.3 ... 69.6 vs 75.9) -- no gain (the kernel is not limited by
// the 3.5 % of redundant reads alone), column-major order is 13 % slower; removed again.  Tile heights 24 .. 48 are
// within 1 % of each other at 3072^2 .. 8192^2 (profiles/r02_t2_rows_sweep.log).
// Software pipelining (profiles/r02_t2_pipe_ab.log): the level-(n+2) update of row j-1 and the level-(n+1) update of
// row j+1 issued as ONE straight-line block between the same two barriers (two independent collision chains per
// warp against the `wait` stalls ncu reports, 102 registers): 82.3 vs 83.2 GLUPS at 16384^2, 73.1 vs 74.9 at 4096^2 in
// an interleaved A/B, bit-identical -- slower, removed again.
#ifndef LBM_T2_COMPACT_RING
#define LBM_T2_COMPACT_RING 1
#endif
#ifndef LBM_T2_MINB
#define LBM_T2_MINB 2
#endif
// LBM_T2_TMA: the nine level-n source segments of a row (258 contiguous elements each, 16-byte aligned supersets of
// the 256 a tile needs) are staged in shared memory by bulk-async copies (cp.async.bulk, SASS UBLKCP) that ONE
// thread issues LBM_T2_STAGES rows ahead and an mbarrier per stage completes: the loads are in flight while the
// CTA's warps collide, without holding registers and without per-thread LDGSTS traffic.  Measured at 16384^2 fp64
// EXACT, 32-row tiles (GLUPS; profiles/r02_t2_variants.log): direct loads, compact ring, 4 CTAs/SM 80.1; TMA 1 stage
// x 4 CTAs/SM 84.0; 2 stages x 3 CTAs 85.1; 3 stages x 2 CTAs 87.5 (shipped: 6.30 TB/s of pass traffic = 0.963 of the
// measured copy peak); 4 stages x 2 CTAs 87.5.  Prefetch depth, not occupancy, hides the DRAM latency here.
#ifndef LBM_T2_TMA
#define LBM_T2_TMA 1
#endif
#ifndef LBM_T2_STAGES
#define LBM_T2_STAGES 3
#endif
// fp32 (16384^2 EXACT, GLUPS at 32- / 64-row tiles, profiles/r02_t2_variants.log): 3 stages x 2 CTAs 143.9 / 147.0,
// 4 x 4 143.8 / 147.2, 6 x 2 128.7 / 131.9, 6 x 3 128.7 / 131.8, 8 x 2 104.7 / 107.6 -- deeper staging does not help:
// at 36 B per cell per step the fp32 kernel is issue-bound (two collisions per cell per pass), not latency-bound.
#ifndef LBM_T2_STAGES_F32
#define LBM_T2_STAGES_F32 3
#endif
#ifndef LBM_T2_MINB_F32
#define LBM_T2_MINB_F32 2
#endif
template <typename T>
__host__ __device__ constexpr int t2_stages() { return sizeof(T) == 4 ? LBM_T2_STAGES_F32 : LBM_T2_STAGES; }
template <typename T>
__host__ __device__ constexpr int t2_minb() { return sizeof(T) == 4 ? LBM_T2_MINB_F32 : LBM_T2_MINB; }

// Shared-memory ring of level-(n+1) rows, one barrier per row.  At iteration j (row j of level n+1 has just been
// written) the level-(n+2) pull of row j-1 reads NW,SW from row j, N,S from row j-1 and NE,SE from row j-2, while
// a warp that is already past the barrier may be writing row j+1: NW,SW need 2 live rows, N,S 3, NE,SE 4.
__host__ __device__ constexpr int ring_slots(int i)
{
    return LBM_T2_COMPACT_RING ? ((i == 6 || i == 7) ? 2 : (i == 2 || i == 4) ? 3 : 4) : 4;
}
// first population-row of population i in the ring (order NW, SW, N, S, NE, SE)
__host__ __device__ constexpr int ring_base(int i)
{
    return i == 6 ? 0 : i == 7 ? ring_slots(6) : i == 2 ? 2 * ring_slots(6) : i == 4 ? 2 * ring_slots(6) + ring_slots(2)
         : i == 5 ? 2 * ring_slots(6) + 2 * ring_slots(2) : 2 * ring_slots(6) + 2 * ring_slots(2) + ring_slots(5);
}
constexpr int T2_RING_ROWS = 2 * ring_slots(6) + 2 * ring_slots(2) + 2 * ring_slots(5);
template <typename T>
__host__ __device__ constexpr int t2_seg_elems() { return T2_TILE + 16 / (int)sizeof(T); }   // staged segment: tile + one 16-byte unit
template <typename T>
__host__ __device__ constexpr int t2_smem_bytes()
{
    return T2_RING_ROWS * T2_TILE * (int)sizeof(T) +
           (LBM_T2_TMA ? t2_stages<T>() * (9 * t2_seg_elems<T>() * (int)sizeof(T) + 16) : 0);
}

// cells closer than w to the perimeter (needs lnx, lny >= 2w)
__host__ __device__ inline long long ring_cells(int lnx, int lny, int w) { return 2ll * w * lny + 2ll * w * (lnx - 2 * w); }

__device__ __forceinline__ void decode_ring(int lnx, int lny, int w, long long t, int &k, int &l)
{
    const long long a = (long long)w * lny;
    if (t < a) { k = (int)(t / lny); l = (int)(t - (long long)k * lny); return; }
    t -= a;
    if (t < a) { const int r = (int)(t / lny); k = lnx - w + r; l = (int)(t - (long long)r * lny); return; }
    t -= a;
    const int m = lnx - 2 * w;
    const long long c = (long long)w * m;
    if (t < c) { l = (int)(t / m); k = w + (int)(t - (long long)l * m); return; }
    t -= c;
    const int r = (int)(t / m);
    l = lny - w + r;
    k = w + (int)(t - (long long)r * m);
}

template <typename T>
__device__ __forceinline__ T *frame_ptr(const FrameView<T> &f, int lnx, int lny, long long pitch, int i, int k, int l)
{
    if (k < FRAME_W) return f.top + ((long long)i * 3 + k) * pitch + (l + PAD_L);
    if (k >= lnx - FRAME_W) return f.bottom + ((long long)i * 3 + (k - (lnx - FRAME_W))) * pitch + (l + PAD_L);
    if (l < FRAME_W) return f.left + ((long long)i * lnx + k) * 4 + l;
    return f.right + ((long long)i * lnx + k) * 4 + (l - (lny - FRAME_W));
}

// ---- sources of a general (boundary-aware) cell update ------------------------------------------------
// level n: the main buffer with its ghost rows and ycol ghost columns (what update_cell<RIM> reads)
template <typename T>
struct SrcMain {
    const StepParams<T> &p;
    const T *src;
    const T *ycol;
    int k, l;
    template <int I>
    __device__ __forceinline__ T pull() const
    {
        if (cy_of(I) == 1 && l == 0) return __ldcg(ycol + (0 * 3 + ycol_slot(I)) * (long long)(p.lnx + 2) + (k - cx_of(I) + 1));
        if (cy_of(I) == -1 && l == p.lny - 1) return __ldcg(ycol + (1 * 3 + ycol_slot(I)) * (long long)(p.lnx + 2) + (k - cx_of(I) + 1));
        return __ldcg(src + (long long)I * p.pop_stride + (long long)(k - cx_of(I) + 1) * p.pitch + (l - cy_of(I) + PAD_L));
    }
    __device__ __forceinline__ T own(int i) const { return __ldcg(src + (long long)i * p.pop_stride + (long long)(k + 1) * p.pitch + (l + PAD_L)); }
};

// level n+1: the frame storage and the level-(n+1) frame ghosts
template <typename T>
struct SrcFrame {
    const StepParams<T> &p;
    FrameView<T> fv;
    int k, l;
    template <int I>
    __device__ __forceinline__ T pull() const
    {
        const int kk = k - cx_of(I), ll = l - cy_of(I);
        if (ll < 0) return __ldcg(fv.gcol + (0 * 3 + ycol_slot(I)) * (long long)(p.lnx + 2) + (kk + 1));
        if (ll >= p.lny) return __ldcg(fv.gcol + (1 * 3 + ycol_slot(I)) * (long long)(p.lnx + 2) + (kk + 1));
        if (kk < 0) return __ldcg(fv.grow + (0 * 3 + xrow_slot(I)) * p.pitch + (ll + PAD_L));
        if (kk >= p.lnx) return __ldcg(fv.grow + (1 * 3 + xrow_slot(I)) * p.pitch + (ll + PAD_L));
        return __ldcg(frame_ptr<T>(fv, p.lnx, p.lny, p.pitch, I, kk, ll));
    }
    __device__ __forceinline__ T own(int i) const { return __ldcg(frame_ptr<T>(fv, p.lnx, p.lny, p.pitch, i, k, l)); }
};

template <typename T, typename Src>
__device__ __forceinline__ void pull9(const Src &s, T (&f)[9])
{
    f[0] = s.template pull<0>();
    f[1] = s.template pull<1>();
    f[2] = s.template pull<2>();
    f[3] = s.template pull<3>();
    f[4] = s.template pull<4>();
    f[5] = s.template pull<5>();
    f[6] = s.template pull<6>();
    f[7] = s.template pull<7>();
    f[8] = s.template pull<8>();
}

// The wall / lid rules of update_cell (cavity_opt2.py:133-177 as a gather) on top of a generic source.
template <typename T, int BC, typename Src>
__device__ __forceinline__ void wall_rules(const StepParams<T> &p, const Src &s, T (&f)[9], int k, int l)
{
    if (BC == BC_PERIODIC) return;
    const long long gk = p.x0 + k, gl = p.y0 + l;
    const bool walls = (BC == BC_CAVITY);
    const bool bottom = (gl == 0), top = (gl == p.gny - 1);
    const bool left = walls && (gk == 0), right = walls && (gk == p.gnx - 1);
    if (!(bottom | top | left | right)) return;
    T lid = T(0), o_nw = T(0), o_ne = T(0);
    if (top) {
        o_nw = s.own(QNW);
        o_ne = s.own(QNE);
        T rho = rn_add(o_nw, s.own(QN));
        rho = rn_add(rho, o_ne);
        rho = rn_add(rho, f[QNW]);
        rho = rn_add(rho, f[QN]);
        rho = rn_add(rho, f[QNE]);
        rho = rn_add(rho, f[QW]);
        rho = rn_add(rho, f[Q0]);
        rho = rn_add(rho, f[QE]);
        const T six_w = rn_mul(T(6), T(1.0 / 36.0));
        lid = rn_mul(rn_mul(six_w, rho), p.u_wall);
    }
#pragma unroll
    for (int i = 1; i < 9; ++i) {
        const bool outside = (cy_of(i) == 1 && bottom) || (cy_of(i) == -1 && top) || (cx_of(i) == 1 && left) || (cx_of(i) == -1 && right);
        if (outside) f[i] = s.own(opp_of(i));
    }
    if (top) {
        if (!left) f[QSE] = rn_add(o_nw, lid);
        if (!right) f[QSW] = rn_sub(o_ne, lid);
    }
}

// Perimeter cell (k, l): store the level-(n+1) populations that leave the block into the neighbours'
// level-(n+1) frame ghosts (mirror of push_halo).
template <typename T>
__device__ __forceinline__ void push_frame_halo(const StepParams<T> &p, int k, int l, const T (&f)[9])
{
    const bool xl = (k == 0), xh = (k == p.lnx - 1), yl = (l == 0), yh = (l == p.lny - 1);
    if (!(xl | xh | yl | yh)) return;
#define LBM_FPUSH_X(D, SIDE, LL, I)                                                              \
    do {                                                                                         \
        const NbrView<T> &nb = p.nbr[D];                                                         \
        const FrameView<T> nf = frame_view<T>(nb.frame, nb.lnx, nb.pitch);                       \
        nf.grow[((SIDE) * 3 + xrow_slot(I)) * nb.pitch + ((LL) + PAD_L)] = f[I];                 \
    } while (0)
#define LBM_FPUSH_Y(D, SIDE, KK, I)                                                              \
    do {                                                                                         \
        const NbrView<T> &nb = p.nbr[D];                                                         \
        const FrameView<T> nf = frame_view<T>(nb.frame, nb.lnx, nb.pitch);                       \
        nf.gcol[((SIDE) * 3 + ycol_slot(I)) * (long long)(nb.lnx + 2) + ((KK) + 1)] = f[I];      \
    } while (0)
    if (xh) { LBM_FPUSH_X(1, 0, l, QE); LBM_FPUSH_X(1, 0, l, QNE); LBM_FPUSH_X(1, 0, l, QSE); }   // -> ghost row -1 of the right neighbour
    if (xl) { LBM_FPUSH_X(0, 1, l, QW); LBM_FPUSH_X(0, 1, l, QNW); LBM_FPUSH_X(0, 1, l, QSW); }   // -> ghost row lnx of the left neighbour
    if (yh) { LBM_FPUSH_Y(3, 0, k, QN); LBM_FPUSH_Y(3, 0, k, QNE); LBM_FPUSH_Y(3, 0, k, QNW); }
    if (yl) { LBM_FPUSH_Y(2, 1, k, QS); LBM_FPUSH_Y(2, 1, k, QSW); LBM_FPUSH_Y(2, 1, k, QSE); }
    if (xh && yh) LBM_FPUSH_Y(7, 0, -1, QNE);
    if (xh && yl) LBM_FPUSH_Y(6, 1, -1, QSE);
    if (xl && yh) LBM_FPUSH_Y(5, 0, p.nbr[5].lnx, QNW);
    if (xl && yl) LBM_FPUSH_Y(4, 1, p.nbr[4].lnx, QSW);
#undef LBM_FPUSH_X
#undef LBM_FPUSH_Y
}

__device__ __forceinline__ void wait_flags(const unsigned long long *flags, unsigned long long want, DevState *st, unsigned long long timeout_ns)
{
    if (threadIdx.x < NUM_DIRS) {
        if (ld_acquire_sys(&flags[threadIdx.x]) < want && *(volatile unsigned int *)&st->error == 0) {
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(&flags[threadIdx.x]) < want) {
                if (global_timer_ns() - t0 > timeout_ns) {
                    atomicExch(&st->error, 1u);
                    break;
                }
            }
        }
    }
    __syncthreads();
}

// One of the two parts of a pass that write the destination buffer (K2, K3) is complete; the second one to get here
// publishes the pass.  The kernels of the next pass are ordered after BOTH parts by the host (stream order / events).
__device__ __forceinline__ void t2_part_done(DevState *st, int par)
{
    if (atomicAdd(&st->pass_done, 1u) == 1u) {
        st->pass_done = 0u;
        const unsigned long long step = *(volatile unsigned long long *)&st->step;
        *(volatile unsigned int *)&st->cur = (unsigned int)(par ^ 1);   // ONE buffer flip per pass, two steps
        *(volatile unsigned long long *)&st->step = step + 2ull;
    }
}

// K1: level n -> n+1 on the frame.
template <typename T, int BC, bool EXACT>
__global__ void __launch_bounds__(TILE_L) t2_frame1_kernel(const __grid_constant__ StepParams<T> p)
{
    DevState *st = p.st;
    const unsigned long long step = *(volatile unsigned long long *)&st->step;
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    wait_flags(st->flag_in, step, st, p.halo_timeout_ns);
    const long long t = (long long)blockIdx.x * TILE_L + threadIdx.x;
    if (t < ring_cells(p.lnx, p.lny, FRAME_W)) {
        int k, l;
        decode_ring(p.lnx, p.lny, FRAME_W, t, k, l);
        const SrcMain<T> s{p, p.buf[par], p.ycol[par], k, l};
        T f[9];
        pull9<T>(s, f);
        wall_rules<T, BC>(p, s, f, k, l);
        d2q9_collide<T, EXACT>(f, p.omega);
        const FrameView<T> fv = frame_view<T>(p.frame, p.lnx, p.pitch);
#pragma unroll
        for (int i = 0; i < 9; ++i) *frame_ptr<T>(fv, p.lnx, p.lny, p.pitch, i, k, l) = f[i];
        push_frame_halo<T>(p, k, l, f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        halo_fence(p.sys_scope);
        const unsigned int prev = atomicAdd(&st->frame_done, 1u);
        if (prev == gridDim.x - 1u) {
            st->frame_done = 0u;
            halo_fence(p.sys_scope);
#pragma unroll
            for (int d = 0; d < NUM_DIRS; ++d) st_relaxed_sys(p.nbr[d].fflag_in + dir_opp(d), step + 1ull);
        }
    }
}

// K3: level n+1 -> n+2 on the cells closer than 2 to the perimeter; completes the double step.
template <typename T, int BC, bool EXACT>
__global__ void __launch_bounds__(TILE_L) t2_frame2_kernel(const __grid_constant__ StepParams<T> p)
{
    DevState *st = p.st;
    const unsigned long long step = *(volatile unsigned long long *)&st->step;
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    T *__restrict__ dst = p.buf[par ^ 1];
    wait_flags(st->fflag_in, step + 1ull, st, p.halo_timeout_ns);
    const long long t = (long long)blockIdx.x * TILE_L + threadIdx.x;
    if (t < ring_cells(p.lnx, p.lny, 2)) {
        int k, l;
        decode_ring(p.lnx, p.lny, 2, t, k, l);
        const SrcFrame<T> s{p, frame_view<T>(p.frame, p.lnx, p.pitch), k, l};
        T f[9];
        pull9<T>(s, f);
        wall_rules<T, BC>(p, s, f, k, l);
        d2q9_collide<T, EXACT>(f, p.omega);
        T *dp = dst + (long long)(k + 1) * p.pitch + (l + PAD_L);
#pragma unroll
        for (int i = 0; i < 9; ++i) dp[(long long)i * p.pop_stride] = f[i];
        push_halo<T>(p, par ^ 1, k, l, f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        halo_fence(p.sys_scope);
        const unsigned int prev = atomicAdd(&st->frame_done, 1u);
        if (prev == gridDim.x - 1u) {
            st->frame_done = 0u;
            halo_fence(p.sys_scope);
#pragma unroll
            for (int d = 0; d < NUM_DIRS; ++d) st_relaxed_sys(p.nbr[d].flag_in + dir_opp(d), step + 2ull);
            t2_part_done(st, par);
        }
    }
}

// K2: level n -> n+2 on the deep interior.  Level-(n+1) rows stream through the tile: the six populations
// that shift in y go through a shared-memory ring (neighbouring threads consume them), the three that do not
// (rest, E, W: consumed by the SAME thread one row later / earlier) stay in registers.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// mbarrier + bulk-async copy (TMA unit, no tensor map: plain 1-D copies)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T, int BC, bool EXACT>
__global__ void __launch_bounds__(T2_TILE, t2_minb<T>()) t2_interior_kernel(const __grid_constant__ StepParams<T> p)
{
    extern __shared__ __align__(128) unsigned char t2_smem_raw[];
    T *ring = reinterpret_cast<T *>(t2_smem_raw);               // [T2_RING_ROWS][T2_TILE]
    const int par = (int)*(volatile unsigned int *)&p.st->cur;
    const T *__restrict__ src = p.buf[par];
    T *__restrict__ dst = p.buf[par ^ 1];

    const int kt = (int)blockIdx.x / p.t2_tiles_l, lt = (int)blockIdx.x - kt * p.t2_tiles_l;
    const int k0 = 2 + kt * p.t2_rows;
    const int k1 = min(k0 + p.t2_rows, p.lnx - 2);
    const int t = threadIdx.x;
    const int lc = lt * T2_W + T2_S - T2_OFF + t;               // this thread's column (level n+1 and level n+2)
    const bool have1 = t >= T2_OFF - 1 && t <= T2_OFF + T2_W && lc >= 1 && lc <= p.lny - 2;   // a level-(n+1) cell at distance >= 1
    const bool have2 = t >= T2_OFF && t < T2_OFF + T2_W && lc >= 2 && lc <= p.lny - 3;        // a level-(n+2) cell at distance >= 2
    const long long row_bytes = p.pitch * (long long)sizeof(T);
    const char *sp = reinterpret_cast<const char *>(src + (long long)k0 * p.pitch + (lc + PAD_L));        // row k0-1
    T *dp = dst + (long long)(k0 + 1) * p.pitch + (lc + PAD_L);                                           // row k0
    T rest_m1 = T(0), e_m1 = T(0), e_m2 = T(0);                 // level n+1: rest of row j-1, E of rows j-1 and j-2
#if LBM_T2_TMA
    constexpr int AL = 16 / (int)sizeof(T), SEG = t2_seg_elems<T>(), NST = t2_stages<T>();
    const T *stage = ring + T2_RING_ROWS * T2_TILE;             // [NST][9][SEG]
    const unsigned stage_s = smem_u32(stage);
    const unsigned full_s = stage_s + NST * 9 * SEG * (unsigned)sizeof(T);   // NST mbarriers, 8 bytes each
    const int lc0 = lt * T2_W + T2_S - T2_OFF;                  // column of thread 0
    // The last tile of a row band is usually narrower than the others (16384 columns: 124 of 254; 4096: 28 of 254):
    // it copies only the 16-byte units its level-(n+1) threads (columns <= lny - 2) read.
    const int ncols = min(T2_TILE, p.lny - 1 - lc0);
    const unsigned seg_bytes = (unsigned)(((ncols + AL - 1 + AL - 1) / AL) * AL * (int)sizeof(T));      // <= SEG elements
    // stage `st` <- the nine source segments of level-(n+1) row `row`: population i from row (row - cx_i), columns
    // from (lc0 - cy_i) rounded down to a 16-byte unit
    auto stage_row = [&](int row, int st) {
        mbar_expect_tx(full_s + 8u * st, 9u * seg_bytes);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const T *g = src + (long long)i * p.pop_stride + (long long)(row - cx_of(i) + 1) * p.pitch + (PAD_L + (((lc0 - cy_of(i)) / AL) * AL));
            bulk_g2s(stage_s + (unsigned)((st * 9 + i) * SEG * (int)sizeof(T)), g, seg_bytes, full_s + 8u * st);
        }
    };
    if (t == 0) {
#pragma unroll
        for (int st = 0; st < NST; ++st) mbar_init(full_s + 8u * st, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
#pragma unroll
        for (int st = 0; st < NST; ++st)
            if (k0 - 1 + st <= k1) stage_row(k0 - 1 + st, st);
    }
    __syncthreads();
    int st_cur = 0;
    unsigned st_par = 0u;
#endif
    int s3 = (k0 - 1) % 3;                                      // ring slot of row j for the 3-slot populations (j & 1, j & 3 for the others)
#pragma unroll 1
    for (int j = k0 - 1; j <= k1; ++j) {
        sp += row_bytes;
        T f[9];
#if LBM_T2_TMA
        mbar_wait(full_s + 8u * st_cur, st_par);
#endif
        if (have1) {
#if LBM_T2_TMA
#pragma unroll
            for (int i = 0; i < 9; ++i) f[i] = stage[(st_cur * 9 + i) * SEG + ((lc0 - cy_of(i)) & (AL - 1)) + t];
#else
            interior_load<T>(p, sp - row_bytes, f);
#endif
            d2q9_collide<T, EXACT>(f, p.omega);
#define LBM_RING_W(I, SLOT) ring[(ring_base(I) + (ring_slots(I) == 2 ? (j & 1) : ring_slots(I) == 3 ? s3 : (j & 3))) * T2_TILE + t] = f[I]
            LBM_RING_W(QN, 0);
            LBM_RING_W(QS, 0);
            LBM_RING_W(QNE, 0);
            LBM_RING_W(QNW, 0);
            LBM_RING_W(QSW, 0);
            LBM_RING_W(QSE, 0);
#undef LBM_RING_W
        }
        __syncthreads();
#if LBM_T2_TMA
        // every thread has read stage st_cur (before the barrier): refill it with the row NST iterations ahead
        if (t == 0 && j + NST <= k1) {
            fence_proxy_async();
            stage_row(j + NST, st_cur);
        }
        if (++st_cur == NST) { st_cur = 0; st_par ^= 1u; }
#endif
        const T rest_0 = f[Q0], e_0 = f[QE], w_0 = f[QW];       // level n+1, row j, this column
        if (j >= k0 + 1) {                                       // level-(n+1) rows j-2, j-1, j are available: emit row j-1
            if (have2) {
                T g[9];
                g[Q0] = rest_m1;                                 // (j-1, lc)
                g[QE] = e_m2;                                    // pulled from row j-2
                g[QW] = w_0;                                     // pulled from row j
                // slot of row j - d for a population with n slots, from the running slot counters of row j
#define LBM_SLOT(I, D) (ring_slots(I) == 2 ? ((j - (D)) & 1) : ring_slots(I) == 3 ? ((s3 + 3 - (D)) % 3) : ((j - (D)) & 3))
#define LBM_RING(I) ring[(ring_base(I) + LBM_SLOT(I, 1 + cx_of(I))) * T2_TILE + (t - cy_of(I))]
                g[QN] = LBM_RING(QN);
                g[QS] = LBM_RING(QS);
                g[QNE] = LBM_RING(QNE);
                g[QNW] = LBM_RING(QNW);
                g[QSW] = LBM_RING(QSW);
                g[QSE] = LBM_RING(QSE);
#undef LBM_RING
#undef LBM_SLOT
                d2q9_collide<T, EXACT>(g, p.omega);
                T *q = dp;
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    *q = g[i];
                    q += p.pop_stride;
                }
            }
            dp += p.pitch;
        }
        e_m2 = e_m1;
        e_m1 = e_0;
        rest_m1 = rest_0;
        s3 = s3 == 2 ? 0 : s3 + 1;
    }
    // the last tile to finish marks the interior part of the pass complete (this CTA read `par` before any flip:
    // the pass is published only after every interior CTA has been counted)
    if (t == 0) {
        const unsigned int prev = atomicAdd(&p.st->t2_done, 1u);
        if (prev == gridDim.x - 1u) {
            p.st->t2_done = 0u;
            t2_part_done(p.st, par);
        }
    }
}

}  // namespace lbm
