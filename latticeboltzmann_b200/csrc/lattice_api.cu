// C ABI of the device-resident lattice (include/lbm_b200.h, group (2)).
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "../../include/lbm_b200.h"
#include "api_common.h"
#include "step_kernel.cuh"
#include "temporal.cuh"
#include "resident.cuh"
#include "aa.cuh"

using namespace lbm;

struct NbrHost {
    bool connected = false;
    char *base = nullptr;   // neighbour's allocation as addressable from this process/device
    lb_export exp{};
};

struct lb_lattice {
    lb_config cfg{};
    size_t elem = 8;
    char *base = nullptr;
    size_t buf_bytes = 0, ycol_off = 0, ycol_bytes = 0, frame_off = 0, state_off = 0, total_bytes = 0;
    long long pitch = 0, pop_stride = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    NbrHost nbr[LB_NUM_DIRS];
    std::map<std::pair<int64_t, uint64_t>, char *> ipc_open;   // (pid, remote base) -> mapped base
    int rows_per_tile = 4;
    unsigned long long halo_timeout_ns = 20ull * 1000 * 1000 * 1000;
    // CUDA graph of GRAPH_STEPS fused steps (launch arguments are step-independent)
    cudaGraphExec_t graph_exec = nullptr;
    cudaStream_t graph_stream = nullptr;
    int graph_rows_per_tile = 0;
    bool use_graph = true;
    int temporal = 0;            // 0: auto (two steps per HBM pass when the block is big enough to fill the GPU), 1: single-step kernel, 2: force temporal blocking
    int t2_rows = 0;             // rows per fused tile; 0: automatic (t2_rows_for)
    // temporal blocking: the frame kernels (K1, K3) run on s_frame concurrently with the fused interior kernel (K2) on `stream`
    cudaStream_t s_frame = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_int = nullptr, ev_frm = nullptr;
    bool overlap_frames = true;
    cudaGraphExec_t graph2_exec = nullptr;   // GRAPH_DOUBLE double steps
    cudaStream_t graph2_stream = nullptr;
    int graph2_rows = 0;
    int64_t steps = 0;
    int cur = 0;                 // host mirror of DevState::cur (buffer holding the current state)
    int64_t launches = 0;
    bool use_resident = true;    // L2-resident single blocks: one cooperative launch for many steps (resident.cuh)
    unsigned long long grid_bar = 0;   // host mirror of DevState::grid_bar
    int resident_ctas = 0;       // co-resident CTAs of the resident kernel on this device (0: not queried yet)
    int resident2_ctas = 0;      // the same for the two-steps-per-barrier variant
    bool use_resident2 = true;
    // shear probe
    void *d_uyk = nullptr, *d_series = nullptr, *d_prod = nullptr;
    int64_t probe_capacity = 0, probe_l_local = -1, probe_step0 = 0;
    // scratch for moments
    void *d_mom = nullptr;
    // slab-pipelined host step
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_done;
    cudaEvent_t ev_tail = nullptr, ev_fin = nullptr;
    // decomposed host step: rim columns staged through pinned memory, begin/run handshake
    void *h_cols = nullptr, *d_cols = nullptr;
    // in-place (AA pattern) lattice: one buffer, layout alternates natural / swapped (aa.cuh)
    bool inplace = false, aa_swapped = false;
    void *d_stash = nullptr, *d_stage = nullptr;
    size_t stage_bytes = 0;
    // per-cell boundary table (LB_SF_TABLE)
    int *d_tab_cells = nullptr;
    long long *d_tab_src = nullptr;
    void *d_tab_add = nullptr;
    unsigned int *d_tab_mask = nullptr;
    int *d_tab_rank = nullptr;
    int tab_n = 0;
    bool host_begun = false;
};

namespace {

int grid_for(long long n, int block);

// Make the lattice's device current for the duration of an API call and restore the caller's
// (the library shares the process with torch, which tracks its own current device).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        err = prev == dev ? cudaSuccess : cudaSetDevice(dev);
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define LBM_ON_DEVICE(L)                    \
    DeviceGuard lbm_guard_((L)->cfg.device); \
    LBM_CUDA(lbm_guard_.err)

void drop_graph(lb_lattice *L)
{
    if (L->graph_exec) cudaGraphExecDestroy(L->graph_exec);
    L->graph_exec = nullptr;
    if (L->graph2_exec) cudaGraphExecDestroy(L->graph2_exec);
    L->graph2_exec = nullptr;
}

DevState *dev_state(lb_lattice *L) { return reinterpret_cast<DevState *>(L->base + L->state_off); }

// Rows per fused tile.  Taller tiles recompute (and re-read) fewer level-(n+1) halo rows -- 2 per tile -- but leave
// fewer tiles to balance over 148 SMs x 2 CTAs.  Two regimes, both measured at 16384^2 fp64 EXACT (GLUPS):
//   burst (fresh lattice, 50 steps, cool GPU = bench.py's timed region; tools/t2_burst_rows.py, profiles/r02_t2_rows_burst.log):
//     32 rows 87.49, 36 rows 87.61, 40 rows 87.76, 48 rows 87.76, 64 rows 87.17, 74 rows 86.59, 96 rows 86.17
//   sustained (interleaved A/B in one process under the power cap; tools/t2_env_ab.py, profiles/r02_t2_rows_ab.log):
//     24 rows 80.2, 32 rows 80.1-81.8, 48 rows 83.0, 64 rows 82.0-83.5, 96 rows 83.0, 128 rows 82.6
// 48 rows is the best of both (the saved arithmetic also saves power).  8192^2 sustained: 32 rows 78.8, 48 rows 79.3, 64
// rows 79.1; 4096^2: 24 .. 48 rows within 1 %, 64 rows -4 % (too few tiles); 2048^2: 48 rows 50.7 vs 60.7 (32 rows) / 60.1
// (16 rows).  fp32 gains 2 % from 64 rows (147.0 vs 143.9).  Smaller lattices take 16-row tiles so that enough tiles
// remain (profiles/r02_t2_rows_sweep.log, r02_t2_small_sweep.log: 1536^2 45.4 (16 rows) vs 39.5 (32 rows) GLUPS).
int t2_rows_for(const lb_lattice *L)
{
    if (L->t2_rows > 0) return L->t2_rows;
    const long long tl = t2_tiles_over(L->cfg.lny);
    auto tiles = [&](int rows) { return tl * ((L->cfg.lnx - 4 + rows - 1) / rows); };
    if (L->cfg.dtype == LB_F32 && tiles(64) >= 1024) return 64;
    if (tiles(48) >= 800) return 48;
    return tiles(32) >= 1024 ? 32 : 16;
}

template <typename T>
StepParams<T> make_params(lb_lattice *L)
{
    StepParams<T> p{};
    p.buf[0] = reinterpret_cast<T *>(L->base);
    p.buf[1] = L->inplace ? p.buf[0] : reinterpret_cast<T *>(L->base + L->buf_bytes);
    p.bc = L->cfg.boundary;
    p.aa_swapped = L->aa_swapped ? 1 : 0;
    p.ycol[0] = reinterpret_cast<T *>(L->base + L->ycol_off);
    p.ycol[1] = reinterpret_cast<T *>(L->base + L->ycol_off + L->ycol_bytes);
    p.frame = reinterpret_cast<T *>(L->base + L->frame_off);
    p.st = dev_state(L);
    p.pop_stride = L->pop_stride;
    p.pitch = L->pitch;
    p.x0 = L->cfg.x0;
    p.y0 = L->cfg.y0;
    p.gnx = L->cfg.gnx;
    p.gny = L->cfg.gny;
    p.lnx = (int)L->cfg.lnx;
    p.lny = (int)L->cfg.lny;
    p.rows_per_tile = L->rows_per_tile;
    p.tiles_l = (p.lny + TILE_L - 1) / TILE_L;
    p.tiles_k = (p.lnx + p.rows_per_tile - 1) / p.rows_per_tile;
    p.all_rim = L->cfg.boundary >= LB_SF_COUETTE ? 1 : 0;
    if (p.all_rim) {
        // simple_flows flavour: wall rules also touch the rows/columns next to the wall layers; on these
        // small (L2-resident, launch-bound) lattices every cell simply takes the general path.
        p.n_perimeter = (long long)p.lnx * p.lny;
        p.tiles_l = p.tiles_k = 0;
    } else {
        p.n_perimeter = (long long)p.lny + (p.lnx > 1 ? p.lny : 0) + (p.lny > 1 ? 2 : 1) * (long long)(p.lnx > 2 ? p.lnx - 2 : 0);
    }
    p.n_rim_ctas = (int)((p.n_perimeter + TILE_L - 1) / TILE_L);
    p.t2_rows = t2_rows_for(L);
    p.t2_tiles_l = t2_tiles_over(p.lny);
    p.t2_tiles_k = p.lnx > 4 ? (p.lnx - 4 + p.t2_rows - 1) / p.t2_rows : 0;
    p.sf_uw6 = (T)((1.0 / 6.0) * L->cfg.u_wall);
    p.rho_in = (T)L->cfg.rho_in;
    p.rho_out = (T)L->cfg.rho_out;
    p.omega = (T)L->cfg.omega;
    p.u_wall = (T)L->cfg.u_wall;
    for (int i = 0; i < 9; ++i) {
        p.ld_off[i] = (long long)sizeof(T) * (i * L->pop_stride - cx_of(i) * L->pitch - cy_of(i));
    }
    p.tab_cells = L->d_tab_cells;
    p.tab_src = L->d_tab_src;
    p.tab_add = static_cast<const T *>(L->d_tab_add);
    p.tab_mask = L->d_tab_mask;
    p.tab_rank = L->d_tab_rank;
    p.tab_n = L->tab_n;
    p.sys_scope = 0;
    p.halo_timeout_ns = L->halo_timeout_ns;
    for (int d = 0; d < LB_NUM_DIRS; ++d) {
        const NbrHost &n = L->nbr[d];
        if (!n.connected) continue;
        if (n.exp.pid != (int64_t)getpid() || n.exp.device != L->cfg.device) p.sys_scope = 1;
        p.nbr[d].buf[0] = reinterpret_cast<T *>(n.base);
        p.nbr[d].buf[1] = reinterpret_cast<T *>(n.base + n.exp.buf_bytes);
        p.nbr[d].ycol[0] = reinterpret_cast<T *>(n.base + n.exp.ycol_offset);
        p.nbr[d].ycol[1] = reinterpret_cast<T *>(n.base + n.exp.ycol_offset + n.exp.ycol_bytes);
        p.nbr[d].flag_in = reinterpret_cast<DevState *>(n.base + n.exp.state_offset)->flag_in;
        p.nbr[d].fflag_in = reinterpret_cast<DevState *>(n.base + n.exp.state_offset)->fflag_in;
        p.nbr[d].frame = reinterpret_cast<T *>(n.base + n.exp.frame_offset);
        p.nbr[d].pop_stride = n.exp.pop_stride;
        p.nbr[d].pitch = n.exp.pitch;
        p.nbr[d].lnx = (int)n.exp.lnx;
        p.nbr[d].lny = (int)n.exp.lny;
    }
    return p;
}

template <typename T, int BC, bool EXACT>
int launch_step_bc(lb_lattice *L, const StepParams<T> &p, bool collide)
{
    const int grid = p.n_rim_ctas + p.tiles_l * p.tiles_k;
    if (collide)
        step_kernel<T, BC, EXACT, true><<<grid, TILE_L, 0, L->stream>>>(p);
    else
        step_kernel<T, BC, EXACT, false><<<grid, TILE_L, 0, L->stream>>>(p);
    return 0;
}

template <typename T>
int launch_step(lb_lattice *L, bool collide)
{
    const StepParams<T> p = make_params<T>(L);
    const bool exact = L->cfg.arith == LB_ARITH_EXACT;
    switch (L->cfg.boundary) {
    case LB_PERIODIC:
        return exact ? launch_step_bc<T, BC_PERIODIC, true>(L, p, collide) : launch_step_bc<T, BC_PERIODIC, false>(L, p, collide);
    case LB_CAVITY:
        return exact ? launch_step_bc<T, BC_CAVITY, true>(L, p, collide) : launch_step_bc<T, BC_CAVITY, false>(L, p, collide);
    case LB_CAVITY_XPERIODIC:
        return exact ? launch_step_bc<T, BC_CAVITY_XPERIODIC, true>(L, p, collide)
                     : launch_step_bc<T, BC_CAVITY_XPERIODIC, false>(L, p, collide);
    case LB_SF_COUETTE:
        return launch_step_bc<T, BC_SF_COUETTE, true>(L, p, collide);
    case LB_SF_POISEUILLE:
        return launch_step_bc<T, BC_SF_POISEUILLE, true>(L, p, collide);
    case LB_SF_SLIDING_LID:
        return launch_step_bc<T, BC_SF_SLIDING_LID, true>(L, p, collide);
    case LB_SF_TABLE:
        return launch_step_bc<T, BC_SF_TABLE, true>(L, p, collide);
    default:
        return lbm_fail(LB_ERR_INVALID, "unknown boundary mode %d", L->cfg.boundary);
    }
}

// One double step (temporal.cuh): frame level n+1, fused deep interior, frame level n+2.
template <typename T, int BC, bool EXACT>
int launch_double_bc(lb_lattice *L, const StepParams<T> &p, int phases, cudaStream_t fs)
{
    // The fused tile's shared memory exceeds the 48 KB default.  Per device and per instantiation; the call is
    // idempotent, so two host threads racing on the bit mask at worst both make it.
    static std::atomic<unsigned long long> attr_done{0ull};
    const unsigned long long dev_bit = 1ull << (L->cfg.device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
        LBM_CUDA(cudaFuncSetAttribute(t2_interior_kernel<T, BC, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, t2_smem_bytes<T>()));
        // The frame kernels share the SMs with the fused tiles (launch_passes): they ask for the same shared-memory
        // carve-out, otherwise an SM has to drain its fused tiles before it can take a frame CTA and again afterwards.
        if (!getenv("LBM_T2_FRAME_CARVEOUT") || atoi(getenv("LBM_T2_FRAME_CARVEOUT"))) {
            LBM_CUDA(cudaFuncSetAttribute(t2_frame1_kernel<T, BC, EXACT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            LBM_CUDA(cudaFuncSetAttribute(t2_frame2_kernel<T, BC, EXACT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    const int g1 = (int)((ring_cells(p.lnx, p.lny, FRAME_W) + TILE_L - 1) / TILE_L);
    const int g3 = (int)((ring_cells(p.lnx, p.lny, 2) + TILE_L - 1) / TILE_L);
    if (phases & 1) t2_frame1_kernel<T, BC, EXACT><<<g1, TILE_L, 0, fs>>>(p);
    if (phases & 2) t2_interior_kernel<T, BC, EXACT><<<p.t2_tiles_l * p.t2_tiles_k, T2_TILE, t2_smem_bytes<T>(), L->stream>>>(p);
    if (phases & 4) t2_frame2_kernel<T, BC, EXACT><<<g3, TILE_L, 0, fs>>>(p);
    return 0;
}

// phases: bit 0 = K1 (frame level n+1), bit 1 = K2 (fused interior), bit 2 = K3 (frame level n+2); the frame kernels
// go to `fs` (default: the lattice's stream), K2 always to the lattice's stream.
template <typename T>
int launch_double(lb_lattice *L, int phases = 7, cudaStream_t fs = nullptr)
{
    const StepParams<T> p = make_params<T>(L);
    const bool exact = L->cfg.arith == LB_ARITH_EXACT;
    if (!fs) fs = L->stream;
    switch (L->cfg.boundary) {
    case LB_PERIODIC:
        return exact ? launch_double_bc<T, BC_PERIODIC, true>(L, p, phases, fs) : launch_double_bc<T, BC_PERIODIC, false>(L, p, phases, fs);
    case LB_CAVITY:
        return exact ? launch_double_bc<T, BC_CAVITY, true>(L, p, phases, fs) : launch_double_bc<T, BC_CAVITY, false>(L, p, phases, fs);
    case LB_CAVITY_XPERIODIC:
        return exact ? launch_double_bc<T, BC_CAVITY_XPERIODIC, true>(L, p, phases, fs) : launch_double_bc<T, BC_CAVITY_XPERIODIC, false>(L, p, phases, fs);
    default:
        return lbm_fail(LB_ERR_INVALID, "temporal blocking supports the periodic and cavity boundaries");
    }
}

int launch_double_any(lb_lattice *L, int phases = 7, cudaStream_t fs = nullptr)
{
    return L->cfg.dtype == LB_F64 ? launch_double<double>(L, phases, fs) : launch_double<float>(L, phases, fs);
}

// The second stream and the events of the overlapped pass (created outside any stream capture).
int ensure_frame_stream(lb_lattice *L)
{
    if (L->s_frame) return 0;                // set last: a partial failure is retried (lb_destroy frees what exists)
    int least = 0, greatest = 0;
    LBM_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    if (!L->ev_fork) LBM_CUDA(cudaEventCreateWithFlags(&L->ev_fork, cudaEventDisableTiming));
    if (!L->ev_int) LBM_CUDA(cudaEventCreateWithFlags(&L->ev_int, cudaEventDisableTiming));
    if (!L->ev_frm) LBM_CUDA(cudaEventCreateWithFlags(&L->ev_frm, cudaEventDisableTiming));
    LBM_CUDA(cudaStreamCreateWithPriority(&L->s_frame, cudaStreamNonBlocking, greatest));   // frame CTAs slip in between the interior's
    return 0;
}

// n passes (2n time steps) of one block.  K1 -> K3 of a pass run on s_frame while K2 runs on the lattice's stream;
// the next pass starts when both are complete: its K2 reads what K3 wrote (cells closer than 2 to the perimeter),
// its K1 reads what K2 wrote (distance 2 and 3) and overwrites the frame storage K3 read.  On return all work is
// joined into the lattice's stream.  Capturable (the graph of GRAPH_DOUBLE passes forks and joins inside).
int launch_passes(lb_lattice *L, int n)
{
    if (!L->overlap_frames) {
        for (int i = 0; i < n; ++i)
            if (int r = launch_double_any(L)) return r;
        return 0;
    }
    LBM_CUDA(cudaEventRecord(L->ev_fork, L->stream));
    LBM_CUDA(cudaStreamWaitEvent(L->s_frame, L->ev_fork, 0));
    for (int i = 0; i < n; ++i) {
        if (int r = launch_double_any(L, 2)) return r;
        LBM_CUDA(cudaEventRecord(L->ev_int, L->stream));
        if (int r = launch_double_any(L, 1 | 4, L->s_frame)) return r;
        LBM_CUDA(cudaEventRecord(L->ev_frm, L->s_frame));
        LBM_CUDA(cudaStreamWaitEvent(L->stream, L->ev_frm, 0));
        LBM_CUDA(cudaStreamWaitEvent(L->s_frame, L->ev_int, 0));
    }
    return 0;
}

constexpr int GRAPH_DOUBLE = 32;     // double steps per graph launch (64 time steps)

int ensure_graph2(lb_lattice *L)
{
    if (L->graph2_exec && L->graph2_stream == L->stream && L->graph2_rows == t2_rows_for(L)) return 0;
    if (L->graph2_exec) cudaGraphExecDestroy(L->graph2_exec);
    L->graph2_exec = nullptr;
    // first launch outside the capture: it sets the kernels' shared-memory attribute
    cudaGraph_t g = nullptr;
    LBM_CUDA(cudaStreamBeginCapture(L->stream, cudaStreamCaptureModeThreadLocal));
    const int r = launch_passes(L, GRAPH_DOUBLE);
    cudaError_t e = cudaStreamEndCapture(L->stream, &g);
    if (r || e != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return r ? r : lbm_fail(LB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(&L->graph2_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        L->graph2_exec = nullptr;
        return lbm_fail(LB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    L->graph2_stream = L->stream;
    L->graph2_rows = t2_rows_for(L);
    return 0;
}

// Fused tiles (counted at 16 rows) a block must offer before the automatic mode prefers temporal blocking: below
// this the deep-interior kernel cannot fill 148 SMs x 2 CTAs and the single-step kernel (graph replay) or the
// resident kernel wins.  Measured crossover (profiles/r02_t2_small_sweep.log, GLUPS two-steps-per-pass vs single
// step): 1024^2 32.5 vs 35.4 (39.5 resident), 1536^2 46.6 vs 38.7, 2048^2 57.7 vs 41.2.
constexpr long long T2_AUTO_MIN_TILES = 512;

bool temporal_ok(const lb_lattice *L)
{
    if (L->temporal == 1 || L->inplace) return false;
    const bool eligible = L->cfg.boundary <= LB_CAVITY_XPERIODIC && L->cfg.lnx >= 16 && L->cfg.lny >= 16 && !L->d_series;
    if (!eligible) return false;
    if (L->temporal == 2) return true;
    const long long tiles = (long long)t2_tiles_over(L->cfg.lny) * ((L->cfg.lnx - 4 + 15) / 16);     // counted in 16-row tiles (decomposition.py)
    return tiles >= T2_AUTO_MIN_TILES;
}

// simple_flows step orders.  The pre-kernels rewrite the current buffer in
// place, so the ghost frame (periodic copies) is refreshed before the pull.
template <typename T>
int launch_sf_prologue(lb_lattice *L)
{
    const StepParams<T> p = make_params<T>(L);
    const long long n = (long long)p.lnx * p.lny;
    if (L->cfg.boundary == LB_SF_COUETTE)
        sf_collide_inplace_kernel<T><<<grid_for(n, 256), 256, 0, L->stream>>>(p);
    else if (L->cfg.boundary == LB_SF_POISEUILLE)
        sf_pressure_kernel<T><<<grid_for(p.lny, 128), 128, 0, L->stream>>>(p);
    else
        return 0;
    halo_refresh_kernel<T><<<grid_for(2ll * (p.lnx + p.lny), 256), 256, 0, L->stream>>>(p, 0, p.lnx);
    L->launches += 2;
    return 0;
}

// ---- resident multi-step kernel (resident.cuh): L2-resident single blocks ---------------------------------
constexpr long long RESIDENT_MAX_CELLS = 1ll << 20;     // 2 x 72 B x 2^20 cells = 151 MB would not stay in L2; above this the per-step kernels win anyway

bool resident_ok(const lb_lattice *L)
{
    if (!L->use_resident || L->inplace || temporal_ok(L)) return false;
    if (L->cfg.lnx * L->cfg.lny > RESIDENT_MAX_CELLS) return false;
    for (int d = 0; d < LB_NUM_DIRS; ++d)
        if (!L->nbr[d].connected || L->nbr[d].base != L->base) return false;   // the wrap is index arithmetic: one self-connected block
    return true;
}

template <typename T, int BC, bool EXACT>
int launch_resident_bc(lb_lattice *L, int64_t nsteps)
{
    auto kernel = resident_kernel<T, BC, EXACT>;
    if (!L->resident_ctas) {
        int per_sm = 0, sms = 0;
        LBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RES_THREADS, 0));
        LBM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, L->cfg.device));
        L->resident_ctas = per_sm * sms;
        if (L->resident_ctas < 2) return lbm_fail(LB_ERR_CUDA, "the resident kernel does not fit the device");
    }
    StepParams<T> p = make_params<T>(L);
    const bool probe = L->d_series != nullptr;
    // Tiny lattices (all 16 x 32 tiles co-resident at one CTA per SM): two steps per grid barrier (resident2_kernel)
    if constexpr (BC <= BC_CAVITY_XPERIODIC) if (L->use_resident2 && nsteps >= 2) {
        auto kernel2 = resident2_kernel<T, BC, EXACT>;
        if (!L->resident2_ctas) {
            int per_sm = 0, sms = 0;
            LBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel2, RES2_THREADS, 0));
            LBM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, L->cfg.device));
            L->resident2_ctas = per_sm * sms > 0 ? per_sm * sms : -1;
        }
        const long long tiles = (long long)((p.lnx + RES2_TX - 1) / RES2_TX) * ((p.lny + RES2_TY - 1) / RES2_TY);
        if (tiles + (probe ? 1 : 0) <= L->resident2_ctas) {
            const int grid2 = (int)tiles + (probe ? 1 : 0);
            while (nsteps >= 2) {
                const int64_t chunk = (nsteps > (1 << 20) ? (1 << 20) : nsteps) & ~(int64_t)1;
                ResidentArgs a{};
                a.nsteps = chunk;
                a.bar_base = L->grid_bar;
                a.probe_l = probe ? (int)L->probe_l_local : -1;
                a.uy_k = L->d_uyk;
                a.series = L->d_series;
                a.capacity = L->probe_capacity;
                a.step0 = (unsigned long long)L->probe_step0;
                a.prod = L->d_prod;
                void *args[] = {&p, &a};
                LBM_CUDA(cudaLaunchCooperativeKernel((void *)kernel2, dim3(grid2), dim3(RES2_THREADS), args, 0, L->stream));
                L->grid_bar += (unsigned long long)grid2 * (unsigned long long)(chunk / 2);
                L->launches++;
                L->steps += chunk;
                L->cur ^= (int)((chunk / 2) & 1);
                nsteps -= chunk;
            }
        }
    }
    const long long n = (long long)p.lnx * p.lny;
    long long workers = (n + RES_THREADS - 1) / RES_THREADS;
    if (workers > L->resident_ctas - (probe ? 1 : 0)) workers = L->resident_ctas - (probe ? 1 : 0);
    const int grid = (int)workers + (probe ? 1 : 0);
    while (nsteps > 0) {
        const int64_t chunk = nsteps > (1 << 20) ? (1 << 20) : nsteps;
        ResidentArgs a{};
        a.nsteps = chunk;
        a.bar_base = L->grid_bar;
        a.couette_shift = BC == BC_SF_COUETTE ? 1 : 0;
        a.probe_l = probe ? (int)L->probe_l_local : -1;
        a.uy_k = L->d_uyk;
        a.series = L->d_series;
        a.capacity = L->probe_capacity;
        a.step0 = (unsigned long long)L->probe_step0;
        a.prod = L->d_prod;
        void *args[] = {&p, &a};
        LBM_CUDA(cudaLaunchCooperativeKernel((void *)kernel, dim3(grid), dim3(RES_THREADS), args, 0, L->stream));
        L->grid_bar += (unsigned long long)grid * (unsigned long long)(chunk + (BC == BC_SF_POISEUILLE ? 1 : 0) + (a.couette_shift ? 1 : 0));
        L->launches++;
        L->steps += chunk;
        L->cur ^= (int)(chunk & 1);
        nsteps -= chunk;
    }
    // the kernel works without the ghost frame: re-establish it for whatever runs next (fused steps, moments, ...)
    halo_refresh_kernel<T><<<grid_for(2ll * (p.lnx + p.lny), 256), 256, 0, L->stream>>>(p, 0, p.lnx);
    L->launches++;
    LBM_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int launch_resident(lb_lattice *L, int64_t nsteps)
{
    const bool exact = L->cfg.arith == LB_ARITH_EXACT;
    switch (L->cfg.boundary) {
    case LB_PERIODIC:
        return exact ? launch_resident_bc<T, BC_PERIODIC, true>(L, nsteps) : launch_resident_bc<T, BC_PERIODIC, false>(L, nsteps);
    case LB_CAVITY:
        return exact ? launch_resident_bc<T, BC_CAVITY, true>(L, nsteps) : launch_resident_bc<T, BC_CAVITY, false>(L, nsteps);
    case LB_CAVITY_XPERIODIC:
        return exact ? launch_resident_bc<T, BC_CAVITY_XPERIODIC, true>(L, nsteps) : launch_resident_bc<T, BC_CAVITY_XPERIODIC, false>(L, nsteps);
    case LB_SF_COUETTE:
        return launch_resident_bc<T, BC_SF_COUETTE, true>(L, nsteps);
    case LB_SF_POISEUILLE:
        return launch_resident_bc<T, BC_SF_POISEUILLE, true>(L, nsteps);
    case LB_SF_SLIDING_LID:
        return launch_resident_bc<T, BC_SF_SLIDING_LID, true>(L, nsteps);
    case LB_SF_TABLE:
        return launch_resident_bc<T, BC_SF_TABLE, true>(L, nsteps);
    default:
        return lbm_fail(LB_ERR_INVALID, "unknown boundary mode %d", L->cfg.boundary);
    }
}

// ---- in-place (AA pattern) stepping, aa.cuh ------------------------------------------------------------------
template <typename T, int BC, bool EXACT>
int launch_aa_bc(lb_lattice *L)
{
    const StepParams<T> p = make_params<T>(L);
    const int grid = p.tiles_l * p.tiles_k;
    const T *stash = static_cast<const T *>(L->d_stash);
    if (BC == BC_CAVITY) {
        aa_corner_stash_kernel<T><<<1, 32, 0, L->stream>>>(p, static_cast<T *>(L->d_stash));
        L->launches++;
    }
    if (L->aa_swapped)
        aa_step_kernel<T, BC, EXACT, true><<<grid, TILE_L, 0, L->stream>>>(p, stash);
    else
        aa_step_kernel<T, BC, EXACT, false><<<grid, TILE_L, 0, L->stream>>>(p, stash);
    L->launches++;
    L->aa_swapped = !L->aa_swapped;
    L->steps++;
    return 0;
}

template <typename T>
int launch_aa(lb_lattice *L)
{
    const bool exact = L->cfg.arith == LB_ARITH_EXACT;
    switch (L->cfg.boundary) {
    case LB_PERIODIC:
        return exact ? launch_aa_bc<T, BC_PERIODIC, true>(L) : launch_aa_bc<T, BC_PERIODIC, false>(L);
    case LB_CAVITY:
        return exact ? launch_aa_bc<T, BC_CAVITY, true>(L) : launch_aa_bc<T, BC_CAVITY, false>(L);
    case LB_CAVITY_XPERIODIC:
        return exact ? launch_aa_bc<T, BC_CAVITY_XPERIODIC, true>(L) : launch_aa_bc<T, BC_CAVITY_XPERIODIC, false>(L);
    default:
        return lbm_fail(LB_ERR_INVALID, "in-place lattices support the periodic and cavity boundaries");
    }
}

// device-side step counter of an in-place lattice (lb_health compares it with the host's)
__global__ void aa_publish_kernel(DevState *st, unsigned long long steps)
{
    st->step = steps;
    for (int d = 0; d < NUM_DIRS; ++d) st->flag_in[d] = steps;
}

int aa_step(lb_lattice *L, int64_t nsteps)
{
    for (int d = 0; d < LB_NUM_DIRS; ++d)
        if (L->nbr[d].base != L->base) return lbm_fail(LB_ERR_STATE, "an in-place lattice is a single self-connected block");
    if (L->d_series) return lbm_fail(LB_ERR_STATE, "the shear probe is not available on in-place lattices");
    for (int64_t s = 0; s < nsteps; ++s)
        if (int r = L->cfg.dtype == LB_F64 ? launch_aa<double>(L) : launch_aa<float>(L)) return r;
    aa_publish_kernel<<<1, 1, 0, L->stream>>>(dev_state(L), (unsigned long long)L->steps);
    LBM_CUDA(cudaGetLastError());
    return 0;
}

int check_ready(lb_lattice *L)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    for (int d = 0; d < LB_NUM_DIRS; ++d)
        if (!L->nbr[d].connected) return lbm_fail(LB_ERR_STATE, "direction slot %d is not connected (lb_connect)", d);
    return 0;
}

int grid_for(long long n, int block) { long long g = (n + block - 1) / block; return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g)); }

}  // namespace

extern "C" {

int lb_abi_version(void) { return LB_ABI_VERSION; }

int lb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lb_create(const lb_config *cfg, lb_lattice **out) { return lb_create_ex(cfg, 0, out); }

int lb_create_ex(const lb_config *cfg, int flags, lb_lattice **out)
{
    if (!cfg || !out) return lbm_fail(LB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->dtype != LB_F32 && cfg->dtype != LB_F64) return lbm_fail(LB_ERR_INVALID, "dtype must be LB_F32 or LB_F64");
    if (cfg->lnx < 1 || cfg->lny < 1 || cfg->gnx < cfg->lnx || cfg->gny < cfg->lny || cfg->x0 < 0 || cfg->y0 < 0 ||
        cfg->x0 + cfg->lnx > cfg->gnx || cfg->y0 + cfg->lny > cfg->gny)
        return lbm_fail(LB_ERR_INVALID, "inconsistent block geometry");
    if (cfg->lnx > (1ll << 30) || cfg->lny > (1ll << 30)) return lbm_fail(LB_ERR_INVALID, "block extent too large");
    if (cfg->boundary < LB_PERIODIC || cfg->boundary > LB_SF_TABLE) return lbm_fail(LB_ERR_INVALID, "unknown boundary");
    if (cfg->arith != LB_ARITH_EXACT && cfg->arith != LB_ARITH_FAST) return lbm_fail(LB_ERR_INVALID, "unknown arith");
    // A wall-bounded box needs two distinct wall rows/columns: with a single row the reference's
    // sequential overwrites (cavity_opt2.py:134-177) alias top and bottom and the gather form does not apply.
    if (cfg->boundary != LB_PERIODIC && (cfg->gny < 2 || (cfg->boundary != LB_CAVITY_XPERIODIC && cfg->gnx < 2)))
        return lbm_fail(LB_ERR_INVALID, "wall-bounded lattices need at least 2 cells across each walled direction");
    if (cfg->boundary >= LB_SF_COUETTE) {
        // simple_flows flavour (config 2): fp64 like the numpy reference, one block, lattice includes the wall layers
        if (cfg->dtype != LB_F64) return lbm_fail(LB_ERR_INVALID, "simple_flows boundaries are fp64 only (the reference is numpy float64)");
        if (cfg->x0 != 0 || cfg->y0 != 0 || cfg->lnx != cfg->gnx || cfg->lny != cfg->gny)
            return lbm_fail(LB_ERR_INVALID, "simple_flows boundaries run on a single block");
        if (cfg->gnx < 3 || cfg->gny < 3) return lbm_fail(LB_ERR_INVALID, "simple_flows lattices (wall layers included) need >= 3 cells per direction");
    }
    if (lb_device_count() <= 0)
        return lbm_fail(LB_ERR_NO_DEVICE, "no CUDA device visible: liblbm_b200 has no CPU fallback");
    if (flags & ~LB_CREATE_INPLACE) return lbm_fail(LB_ERR_INVALID, "unknown creation flags 0x%x", flags);
    if ((flags & LB_CREATE_INPLACE) && (cfg->boundary > LB_CAVITY_XPERIODIC || cfg->lnx != cfg->gnx || cfg->lny != cfg->gny))
        return lbm_fail(LB_ERR_INVALID, "in-place lattices: one block, periodic or cavity boundaries");
    lb_lattice *L = new lb_lattice();
    L->cfg = *cfg;
    L->inplace = (flags & LB_CREATE_INPLACE) != 0;
    DeviceGuard guard(cfg->device);
    if (guard.err != cudaSuccess) {
        delete L;
        return lbm_fail(LB_ERR_CUDA, "cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(guard.err));
    }
    L->elem = cfg->dtype == LB_F64 ? 8 : 4;
    // pitch: multiple of 32 elements (>= 128 B), room for PAD_L, lny cells and the upper ghost column.
    L->pitch = ((cfg->lny + PAD_L + 1 + 31) / 32) * 32;
    L->pop_stride = (cfg->lnx + 2) * L->pitch;
    L->buf_bytes = (size_t)9 * L->pop_stride * L->elem;
    L->buf_bytes = (L->buf_bytes + 255) / 256 * 256;
    L->ycol_off = (L->inplace ? 1 : 2) * L->buf_bytes;
    L->ycol_bytes = ((size_t)6 * (cfg->lnx + 2) * L->elem + 255) / 256 * 256;   // one ghost-column array per buffer
    L->frame_off = L->ycol_off + 2 * L->ycol_bytes;
    const size_t frame_bytes = ((size_t)frame_elems(cfg->lnx, L->pitch) * L->elem + 255) / 256 * 256;
    L->state_off = L->frame_off + frame_bytes;
    L->total_bytes = L->state_off + 256;
    cudaError_t e = cudaMalloc(&L->base, L->total_bytes);
    if (e != cudaSuccess) {
        const size_t want = L->total_bytes;
        delete L;
        cudaGetLastError();
        return lbm_fail(LB_ERR_CUDA, "cudaMalloc(%zu bytes for the f buffer(s)): %s", want, cudaGetErrorString(e));
    }
    e = cudaStreamCreateWithFlags(&L->own_stream, cudaStreamNonBlocking);
    L->stream = L->own_stream;
    if (e == cudaSuccess) e = cudaEventCreate(&L->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&L->ev1);
    if (e == cudaSuccess) e = cudaMemsetAsync(L->base, 0, L->total_bytes, L->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        lb_destroy(L);
        return lbm_fail(LB_ERR_CUDA, "lattice set-up failed: %s", cudaGetErrorString(e));
    }
    if (L->inplace) {
        if (cudaMalloc(&L->d_stash, 4 * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            lb_destroy(L);
            return lbm_fail(LB_ERR_CUDA, "lattice set-up failed: stash allocation");
        }
        L->temporal = 1;
        L->use_resident = false;
    }
    L->rows_per_tile = cfg->dtype == LB_F64 ? 4 : 8;     // measured optima
    if (const char *t = getenv("LBM_TEMPORAL")) L->temporal = (atoi(t) >= 0 && atoi(t) <= 2) ? atoi(t) : 0;
    if (const char *t = getenv("LBM_RESIDENT")) L->use_resident = atoi(t) != 0;
    if (const char *t = getenv("LBM_RESIDENT2")) L->use_resident2 = atoi(t) != 0;
    if (const char *t = getenv("LBM_T2_ROWS")) if (atoi(t) > 0) L->t2_rows = atoi(t);
    if (const char *t = getenv("LBM_T2_OVERLAP")) L->overlap_frames = atoi(t) != 0;
    const char *env = getenv("LBM_ROWS_PER_TILE");
    if (env && atoi(env) > 0) L->rows_per_tile = atoi(env);
    *out = L;
    return 0;
}

int lb_destroy(lb_lattice *L)
{
    if (!L) return 0;
    DeviceGuard guard(L->cfg.device);
    // L->stream may be borrowed (another block's or torch's) and already gone: sync the device.
    cudaDeviceSynchronize();
    drop_graph(L);
    for (auto &kv : L->ipc_open) cudaIpcCloseMemHandle(kv.second);
    if (L->d_uyk) cudaFree(L->d_uyk);
    if (L->d_series) cudaFree(L->d_series);
    if (L->d_prod) cudaFree(L->d_prod);
    if (L->d_mom) cudaFree(L->d_mom);
    for (auto e : L->ev_up) cudaEventDestroy(e);
    for (auto e : L->ev_done) cudaEventDestroy(e);
    if (L->ev_tail) cudaEventDestroy(L->ev_tail);
    if (L->ev_fin) cudaEventDestroy(L->ev_fin);
    if (L->d_tab_cells) cudaFree(L->d_tab_cells);
    if (L->d_tab_src) cudaFree(L->d_tab_src);
    if (L->d_tab_add) cudaFree(L->d_tab_add);
    if (L->d_tab_mask) cudaFree(L->d_tab_mask);
    if (L->d_tab_rank) cudaFree(L->d_tab_rank);
    if (L->d_stash) cudaFree(L->d_stash);
    if (L->d_stage) cudaFree(L->d_stage);
    if (L->h_cols) cudaFreeHost(L->h_cols);
    if (L->d_cols) cudaFree(L->d_cols);
    if (L->s_h2d) cudaStreamDestroy(L->s_h2d);
    if (L->s_d2h) cudaStreamDestroy(L->s_d2h);
    if (L->ev_fork) cudaEventDestroy(L->ev_fork);
    if (L->ev_int) cudaEventDestroy(L->ev_int);
    if (L->ev_frm) cudaEventDestroy(L->ev_frm);
    if (L->s_frame) cudaStreamDestroy(L->s_frame);
    if (L->ev0) cudaEventDestroy(L->ev0);
    if (L->ev1) cudaEventDestroy(L->ev1);
    if (L->own_stream) cudaStreamDestroy(L->own_stream);
    if (L->base) cudaFree(L->base);
    delete L;
    return 0;
}

int lb_set_stream(lb_lattice *L, void *s)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    L->stream = s ? reinterpret_cast<cudaStream_t>(s) : L->own_stream;
    return 0;
}

void *lb_get_stream(lb_lattice *L) { return L ? reinterpret_cast<void *>(L->stream) : nullptr; }

int lb_set_rows_per_tile(lb_lattice *L, int rows)
{
    if (!L || rows < 1) return lbm_fail(LB_ERR_INVALID, "rows_per_tile must be >= 1");
    L->rows_per_tile = rows;
    return 0;
}

int lb_set_halo_timeout_ms(lb_lattice *L, int64_t ms)
{
    if (!L || ms < 1) return lbm_fail(LB_ERR_INVALID, "timeout must be >= 1 ms");
    L->halo_timeout_ns = (unsigned long long)ms * 1000000ull;
    drop_graph(L);      // the timeout is a launch argument
    return 0;
}

/* One phase of a double step (1: frame level n+1, 2: fused deep interior, 3: frame level n+2, which
 * completes the step).  For drivers that run several blocks on ONE stream: the phases of all blocks must be
 * interleaved (all 1s, all 2s, all 3s), because phase 3 of a block waits for phase 1 of its neighbours. */
int lb_double_step_phase(lb_lattice *L, int phase)
{
    if (int r = check_ready(L)) return r;
    if (phase < 1 || phase > 3) return lbm_fail(LB_ERR_INVALID, "phase must be 1, 2 or 3");
    if (!temporal_ok(L)) return lbm_fail(LB_ERR_STATE, "temporal blocking is not available for this lattice");
    LBM_ON_DEVICE(L);
    int r = launch_double_any(L, 1 << (phase - 1));
    if (r) return r;
    L->launches++;
    if (phase == 3) {
        L->steps += 2;
        L->cur ^= 1;
    }
    LBM_CUDA(cudaGetLastError());
    return 0;
}

/* Rows per fused tile that lb_step would use for this lattice in two-steps-per-pass mode. */
int lb_temporal_rows(lb_lattice *L) { return L ? t2_rows_for(L) : 0; }

/* 1 if lb_step advances this lattice two steps per pass (temporal blocking), else 0. */
int lb_temporal_active(lb_lattice *L) { return L && temporal_ok(L) ? 1 : 0; }

int lb_set_temporal(lb_lattice *L, int steps_per_pass, int rows_per_tile)
{
    if (!L || steps_per_pass < 0 || steps_per_pass > 2) return lbm_fail(LB_ERR_INVALID, "steps_per_pass must be 0 (auto), 1 or 2");
    L->temporal = steps_per_pass;
    if (rows_per_tile > 0) L->t2_rows = rows_per_tile;
    return 0;
}

int lb_set_use_graph(lb_lattice *L, int on)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    L->use_graph = on != 0;
    return 0;
}

int lb_set_boundary_table(lb_lattice *L, int64_t n, const int64_t *cells, const int64_t *src, const double *add)
{
    if (!L || n < 0 || (n > 0 && (!cells || !src || !add))) return lbm_fail(LB_ERR_INVALID, "bad argument");
    if (L->cfg.boundary != LB_SF_TABLE) return lbm_fail(LB_ERR_STATE, "the lattice was not created with LB_SF_TABLE");
    const int64_t lnx = L->cfg.lnx, lny = L->cfg.lny, ncell = lnx * lny;
    if (n > ncell) return lbm_fail(LB_ERR_INVALID, "more table cells than lattice cells");
    // entries sorted by cell, so that a cell's table index is its rank among the set bits of the mask
    std::vector<int64_t> order((size_t)n);
    for (int64_t j = 0; j < n; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return cells[x] < cells[y]; });
    std::vector<int> h_cells((size_t)n);
    std::vector<long long> h_src((size_t)n * 9);
    std::vector<double> h_add((size_t)n * 9);
    std::vector<unsigned int> h_mask((size_t)(ncell + 31) / 32, 0u);
    for (int64_t jj = 0; jj < n; ++jj) {
        const int64_t j = order[jj];
        if (cells[j] < 0 || cells[j] >= ncell) return lbm_fail(LB_ERR_INVALID, "table cell %lld is outside the lattice", (long long)cells[j]);
        if ((h_mask[cells[j] >> 5] >> (cells[j] & 31)) & 1u) return lbm_fail(LB_ERR_INVALID, "table cell %lld is listed twice", (long long)cells[j]);
        h_mask[cells[j] >> 5] |= 1u << (cells[j] & 31);
        h_cells[jj] = (int)cells[j];
        for (int i = 0; i < 9; ++i) {
            const int64_t e = src[9 * j + i];
            if (e < 0 || e >= 9 * ncell) return lbm_fail(LB_ERR_INVALID, "table source %lld is outside the state", (long long)e);
            const int64_t ch = e / ncell, k = (e - ch * ncell) / lny, l = e - ch * ncell - k * lny;
            h_src[9 * jj + i] = ch * L->pop_stride + (k + 1) * L->pitch + l + PAD_L;      // element offset inside a buffer
            h_add[9 * jj + i] = add[9 * j + i];
        }
    }
    std::vector<int> h_rank(h_mask.size());
    int running = 0;
    for (size_t w = 0; w < h_mask.size(); ++w) {
        h_rank[w] = running;
        running += __builtin_popcount(h_mask[w]);
    }
    LBM_ON_DEVICE(L);
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    if (L->d_tab_cells) cudaFree(L->d_tab_cells);
    if (L->d_tab_src) cudaFree(L->d_tab_src);
    if (L->d_tab_add) cudaFree(L->d_tab_add);
    if (L->d_tab_mask) cudaFree(L->d_tab_mask);
    if (L->d_tab_rank) cudaFree(L->d_tab_rank);
    L->d_tab_cells = nullptr; L->d_tab_src = nullptr; L->d_tab_add = nullptr; L->d_tab_mask = nullptr; L->d_tab_rank = nullptr;
    L->tab_n = 0;
    LBM_CUDA(cudaMalloc(&L->d_tab_mask, h_mask.size() * sizeof(unsigned int)));
    LBM_CUDA(cudaMemcpy(L->d_tab_mask, h_mask.data(), h_mask.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    LBM_CUDA(cudaMalloc(&L->d_tab_rank, h_rank.size() * sizeof(int)));
    LBM_CUDA(cudaMemcpy(L->d_tab_rank, h_rank.data(), h_rank.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (n > 0) {
        LBM_CUDA(cudaMalloc(&L->d_tab_cells, (size_t)n * sizeof(int)));
        LBM_CUDA(cudaMalloc(&L->d_tab_src, (size_t)n * 9 * sizeof(long long)));
        LBM_CUDA(cudaMalloc(&L->d_tab_add, (size_t)n * 9 * sizeof(double)));
        LBM_CUDA(cudaMemcpy(L->d_tab_cells, h_cells.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        LBM_CUDA(cudaMemcpy(L->d_tab_src, h_src.data(), (size_t)n * 9 * sizeof(long long), cudaMemcpyHostToDevice));
        LBM_CUDA(cudaMemcpy(L->d_tab_add, h_add.data(), (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice));
    }
    L->tab_n = (int)n;
    return 0;
}

int lb_set_resident(lb_lattice *L, int on)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    L->use_resident = on != 0;
    return 0;
}

int lb_sync(lb_lattice *L)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    LBM_ON_DEVICE(L);
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    return 0;
}

int lb_get_export(lb_lattice *L, lb_export *out)
{
    if (!L || !out) return lbm_fail(LB_ERR_INVALID, "null argument");
    memset(out, 0, sizeof(*out));
    LBM_ON_DEVICE(L);
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaError_t e = cudaIpcGetMemHandle(&h, L->base);
    if (e == cudaSuccess)
        memcpy(out->ipc_mem_handle, &h, 64);
    else
        cudaGetLastError();   // IPC unsupported here: same-process wiring still works
    out->local_base = reinterpret_cast<uint64_t>(L->base);
    out->pid = (int64_t)getpid();
    out->device = L->cfg.device;
    out->dtype = L->cfg.dtype;
    out->lnx = L->cfg.lnx;
    out->lny = L->cfg.lny;
    out->pitch = L->pitch;
    out->pop_stride = L->pop_stride;
    out->buf_bytes = (int64_t)L->buf_bytes;
    out->ycol_offset = (int64_t)L->ycol_off;
    out->ycol_bytes = (int64_t)L->ycol_bytes;
    out->frame_offset = (int64_t)L->frame_off;
    out->state_offset = (int64_t)L->state_off;
    out->total_bytes = (int64_t)L->total_bytes;
    return 0;
}

int lb_connect(lb_lattice *L, int dir, const lb_export *nb)
{
    if (!L || !nb || dir < 0 || dir >= LB_NUM_DIRS) return lbm_fail(LB_ERR_INVALID, "bad argument");
    if (nb->dtype != L->cfg.dtype) return lbm_fail(LB_ERR_INVALID, "neighbour dtype differs");
    // faces must match: x-neighbours share lny, y-neighbours share lnx
    if (dir_dy(dir) == 0 && nb->lny != L->cfg.lny) return lbm_fail(LB_ERR_INVALID, "x-neighbour with different lny");
    if (dir_dx(dir) == 0 && nb->lnx != L->cfg.lnx) return lbm_fail(LB_ERR_INVALID, "y-neighbour with different lnx");
    LBM_ON_DEVICE(L);
    char *base = nullptr;
    if (nb->pid == (int64_t)getpid()) {
        base = reinterpret_cast<char *>(nb->local_base);
        if (nb->device != L->cfg.device) {
            int can = 0;
            LBM_CUDA(cudaDeviceCanAccessPeer(&can, L->cfg.device, nb->device));
            if (!can) return lbm_fail(LB_ERR_CUDA, "device %d cannot access peer %d", L->cfg.device, nb->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return lbm_fail(LB_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
    } else {
        auto key = std::make_pair(nb->pid, nb->local_base);
        auto it = L->ipc_open.find(key);
        if (it != L->ipc_open.end()) {
            base = it->second;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, nb->ipc_mem_handle, 64);
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return lbm_fail(LB_ERR_CUDA, "cudaIpcOpenMemHandle(pid %lld): %s", (long long)nb->pid, cudaGetErrorString(e));
            base = static_cast<char *>(ptr);
            L->ipc_open[key] = base;
        }
    }
    drop_graph(L);
    L->nbr[dir].connected = true;
    L->nbr[dir].base = base;
    L->nbr[dir].exp = *nb;
    return 0;
}

int lb_halo_refresh(lb_lattice *L)
{
    if (int r = check_ready(L)) return r;
    LBM_ON_DEVICE(L);
    const long long n_rim = 2 * (L->cfg.lnx + L->cfg.lny);
    if (L->cfg.dtype == LB_F64)
        halo_refresh_kernel<double><<<grid_for(n_rim, 256), 256, 0, L->stream>>>(make_params<double>(L), 0, (int)L->cfg.lnx);
    else
        halo_refresh_kernel<float><<<grid_for(n_rim, 256), 256, 0, L->stream>>>(make_params<float>(L), 0, (int)L->cfg.lnx);
    LBM_CUDA(cudaGetLastError());
    L->launches++;
    return 0;
}

// Download of rows [k_lo, k_hi) while an in-place lattice is in the swapped layout: gathered into natural order
// slab by slab through a device staging array.
static int aa_download_rows(lb_lattice *L, void *host, int64_t k_lo, int64_t k_hi)
{
    const size_t e = L->elem, row = (size_t)L->cfg.lny * e;
    const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(k_hi - k_lo, (int64_t)((256u << 20) / (9 * row))));
    const size_t need = (size_t)9 * slab * row;
    if (L->stage_bytes < need) {
        if (L->d_stage) cudaFree(L->d_stage);
        L->d_stage = nullptr;
        L->stage_bytes = 0;
        LBM_CUDA(cudaMalloc(&L->d_stage, need));
        L->stage_bytes = need;
    }
    const int64_t rows_total = k_hi - k_lo;
    for (int64_t a = k_lo; a < k_hi; a += slab) {
        const int64_t b = std::min(k_hi, a + slab), n = (b - a) * L->cfg.lny;
        if (L->cfg.dtype == LB_F64)
            aa_gather_rows_kernel<double><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<double>(L), (double *)L->d_stage, (int)a, (int)b);
        else
            aa_gather_rows_kernel<float><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<float>(L), (float *)L->d_stage, (int)a, (int)b);
        LBM_CUDA(cudaGetLastError());
        L->launches++;
        for (int i = 0; i < 9; ++i)
            LBM_CUDA(cudaMemcpyAsync(static_cast<char *>(host) + ((size_t)i * rows_total + (size_t)(a - k_lo)) * row,
                                     static_cast<char *>(L->d_stage) + (size_t)i * (size_t)n * e, (size_t)n * e, cudaMemcpyDeviceToHost, L->stream));
        LBM_CUDA(cudaStreamSynchronize(L->stream));      // the staging array is reused by the next slab
    }
    return 0;
}

static int copy_f(lb_lattice *L, void *host, bool upload)
{
    if (!L || !host) return lbm_fail(LB_ERR_INVALID, "null argument");
    LBM_ON_DEVICE(L);
    if (L->inplace && L->aa_swapped) {
        if (!upload) return aa_download_rows(L, host, 0, L->cfg.lnx);
        L->aa_swapped = false;      // a whole-lattice upload (re)starts in the natural layout
    }
    const size_t e = L->elem;
    char *cur = L->base + (size_t)L->cur * L->buf_bytes;
    const size_t row = (size_t)L->cfg.lny * e;
    for (int i = 0; i < 9; ++i) {
        char *d = cur + ((size_t)i * L->pop_stride + (size_t)L->pitch + PAD_L) * e;
        char *h = static_cast<char *>(host) + (size_t)i * L->cfg.lnx * row;
        if (upload)
            LBM_CUDA(cudaMemcpy2DAsync(d, (size_t)L->pitch * e, h, row, row, (size_t)L->cfg.lnx, cudaMemcpyHostToDevice, L->stream));
        else
            LBM_CUDA(cudaMemcpy2DAsync(h, row, d, (size_t)L->pitch * e, row, (size_t)L->cfg.lnx, cudaMemcpyDeviceToHost, L->stream));
    }
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    return 0;
}

int lb_upload_f(lb_lattice *L, const void *host_f) { return copy_f(L, const_cast<void *>(host_f), true); }
int lb_download_f(lb_lattice *L, void *host_f) { return copy_f(L, host_f, false); }

int lb_init_equilibrium(lb_lattice *L, const void *rho, const void *ux, const void *uy)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    LBM_ON_DEVICE(L);
    L->aa_swapped = false;          // in-place lattices restart in the natural layout
    const long long n = L->cfg.lnx * L->cfg.lny;
    const size_t bytes = (size_t)n * L->elem;
    struct Tmp {      // freed on every return path
        void *d[3] = {nullptr, nullptr, nullptr};
        ~Tmp() { for (void *q : d) if (q) cudaFree(q); }
    } tmp;
    void **d = tmp.d;
    const void *h[3] = {rho, ux, uy};
    for (int j = 0; j < 3; ++j)
        if (h[j]) {
            LBM_CUDA(cudaMalloc(&d[j], bytes));
            LBM_CUDA(cudaMemcpyAsync(d[j], h[j], bytes, cudaMemcpyHostToDevice, L->stream));
        }
    if (L->cfg.dtype == LB_F64)
        init_equilibrium_kernel<double><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<double>(L), (const double *)d[0], (const double *)d[1], (const double *)d[2]);
    else
        init_equilibrium_kernel<float><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<float>(L), (const float *)d[0], (const float *)d[1], (const float *)d[2]);
    LBM_CUDA(cudaGetLastError());
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    L->launches++;
    return 0;
}

namespace {
constexpr int GRAPH_STEPS = 64;

// One launch of the instantiated graph = GRAPH_STEPS fused steps.  Valid because the kernel reads the
// step number (buffer parity, flag values) from device memory: every launch has identical arguments.
int ensure_graph(lb_lattice *L)
{
    if (L->graph_exec && L->graph_stream == L->stream && L->graph_rows_per_tile == L->rows_per_tile) return 0;
    drop_graph(L);
    cudaGraph_t g = nullptr;
    LBM_CUDA(cudaStreamBeginCapture(L->stream, cudaStreamCaptureModeThreadLocal));
    int r = 0;
    for (int s = 0; s < GRAPH_STEPS && !r; ++s) r = L->cfg.dtype == LB_F64 ? launch_step<double>(L, true) : launch_step<float>(L, true);
    cudaError_t e = cudaStreamEndCapture(L->stream, &g);
    if (r || e != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return r ? r : lbm_fail(LB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(&L->graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        L->graph_exec = nullptr;
        return lbm_fail(LB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    L->graph_stream = L->stream;
    L->graph_rows_per_tile = L->rows_per_tile;
    return 0;
}
}  // namespace

int lb_step(lb_lattice *L, int64_t nsteps)
{
    if (int r = check_ready(L)) return r;
    if (nsteps < 0) return lbm_fail(LB_ERR_INVALID, "nsteps < 0");
    LBM_ON_DEVICE(L);
    const bool sf = L->cfg.boundary >= LB_SF_COUETTE;
    if (L->inplace) return aa_step(L, nsteps);
    // Temporal blocking: two steps per pass over HBM (three launches per double step).
    if (temporal_ok(L)) {
        if (L->overlap_frames)
            if (int r = ensure_frame_stream(L)) return r;
        if (L->use_graph && nsteps >= 4 * GRAPH_DOUBLE) {
            // one eager double step first: it sets the fused kernel's shared-memory attribute outside any capture
            int r = launch_passes(L, 1);
            if (r) return r;
            L->launches += 3;
            L->steps += 2;
            L->cur ^= 1;
            nsteps -= 2;
            if ((r = ensure_graph2(L))) return r;
            while (nsteps >= 2 * GRAPH_DOUBLE) {
                LBM_CUDA(cudaGraphLaunch(L->graph2_exec, L->stream));
                L->launches += 3 * GRAPH_DOUBLE;
                L->steps += 2 * GRAPH_DOUBLE;          // GRAPH_DOUBLE buffer flips: an even number
                nsteps -= 2 * GRAPH_DOUBLE;
            }
        }
        if (nsteps >= 2) {
            const int passes = (int)(nsteps / 2);
            if (int r = launch_passes(L, passes)) return r;
            L->launches += 3ll * passes;
            L->steps += 2ll * passes;
            L->cur ^= passes & 1;
            nsteps -= 2ll * passes;
        }
        LBM_CUDA(cudaGetLastError());
    }
    // L2-resident single blocks: ONE cooperative launch for all remaining steps (probe and simple_flows step orders inside).
    if (nsteps >= 2 && resident_ok(L)) {
        int r = L->cfg.dtype == LB_F64 ? launch_resident<double>(L, nsteps) : (sf ? lbm_fail(LB_ERR_INVALID, "simple_flows is fp64") : launch_resident<float>(L, nsteps));
        if (r) return r;
        nsteps = 0;
    }
    // Long runs of plain fused steps replay a CUDA graph (one host call per GRAPH_STEPS launches).
    if (L->use_graph && !sf && !L->d_series && nsteps >= 2 * GRAPH_STEPS) {
        if (int r = ensure_graph(L)) return r;
        while (nsteps >= GRAPH_STEPS) {
            LBM_CUDA(cudaGraphLaunch(L->graph_exec, L->stream));
            L->launches += GRAPH_STEPS;
            L->steps += GRAPH_STEPS;          // an even number of buffer flips
            nsteps -= GRAPH_STEPS;
        }
    }
    for (int64_t s = 0; s < nsteps; ++s) {
        if (sf) launch_sf_prologue<double>(L);
        // Couette collides BEFORE streaming (prologue), so its fused pass streams + reflects only.
        const bool collide = L->cfg.boundary != LB_SF_COUETTE;
        int r = L->cfg.dtype == LB_F64 ? launch_step<double>(L, collide) : launch_step<float>(L, collide);
        if (r) return r;
        L->launches++;
        L->steps++;
        L->cur ^= 1;
        if (L->cfg.boundary == LB_SF_TABLE && L->tab_n > 0) {
            // the listed cells again, from the previous buffer with their table; then the ghosts of the rim cells among them
            const StepParams<double> p = make_params<double>(L);
            table_cells_kernel<double><<<grid_for(L->tab_n, 128), 128, 0, L->stream>>>(p);
            halo_refresh_kernel<double><<<grid_for(2ll * (p.lnx + p.lny), 256), 256, 0, L->stream>>>(p, 0, p.lnx);
            L->launches += 2;
        }
        if (L->d_series) {
            if (L->cfg.dtype == LB_F64)
                shear_probe_kernel<double><<<1, 256, 0, L->stream>>>(make_params<double>(L), (int)L->probe_l_local, (const double *)L->d_uyk,
                                                                      (double *)L->d_series, L->probe_capacity, (unsigned long long)L->probe_step0);
            else
                shear_probe_kernel<float><<<1, 256, 0, L->stream>>>(make_params<float>(L), (int)L->probe_l_local, (const float *)L->d_uyk,
                                                                     (float *)L->d_series, L->probe_capacity, (unsigned long long)L->probe_step0);
            L->launches++;
        }
    }
    LBM_CUDA(cudaGetLastError());
    return 0;
}

/* Stream + boundary handling only (no collision): PyLB.stream /
 * stream_and_bounce_back as a stand-alone operation (cavity_opt2.py:109-177). */
int lb_stream_only(lb_lattice *L, int64_t nsteps)
{
    if (int r = check_ready(L)) return r;
    if (L->inplace) return lbm_fail(LB_ERR_STATE, "lb_stream_only is not available on in-place lattices");
    LBM_ON_DEVICE(L);
    for (int64_t s = 0; s < nsteps; ++s) {
        int r = L->cfg.dtype == LB_F64 ? launch_step<double>(L, false) : launch_step<float>(L, false);
        if (r) return r;
        L->launches++;
        L->steps++;
        L->cur ^= 1;
    }
    LBM_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

namespace {

template <typename T>
int launch_rows(lb_lattice *L, const StepParams<T> &p, int k_lo, int k_hi)
{
    const int tiles_l = (p.lny + TILE_L - 1) / TILE_L;
    const int grid = (k_hi - k_lo) * tiles_l;
    const bool exact = L->cfg.arith == LB_ARITH_EXACT;
#define LBM_ROWS(BC)                                                                                \
    (exact ? step_rows_kernel<T, BC, true><<<grid, TILE_L, 0, L->stream>>>(p, k_lo, k_hi)            \
           : step_rows_kernel<T, BC, false><<<grid, TILE_L, 0, L->stream>>>(p, k_lo, k_hi))
    switch (L->cfg.boundary) {
    case LB_PERIODIC: LBM_ROWS(BC_PERIODIC); break;
    case LB_CAVITY: LBM_ROWS(BC_CAVITY); break;
    case LB_CAVITY_XPERIODIC: LBM_ROWS(BC_CAVITY_XPERIODIC); break;
    default: return lbm_fail(LB_ERR_INVALID, "lb_step_host supports the periodic and cavity boundaries");
    }
#undef LBM_ROWS
    return 0;
}

// rows [k_lo, k_hi) of all 9 populations between buffer `par` and a host array (9, hrows, lny) whose first row
// is lattice row hk0 (the whole block by default: hk0 = 0, hrows = lnx)
int copy_rows(lb_lattice *L, void *host, int par, int64_t k_lo, int64_t k_hi, bool upload, cudaStream_t st,
              int64_t hk0 = 0, int64_t hrows = -1)
{
    const size_t e = L->elem, row = (size_t)L->cfg.lny * e;
    if (hrows < 0) hrows = L->cfg.lnx;
    char *buf = L->base + (size_t)par * L->buf_bytes;
    for (int i = 0; i < 9; ++i) {
        char *d = buf + ((size_t)i * L->pop_stride + (size_t)(k_lo + 1) * L->pitch + PAD_L) * e;
        char *h = static_cast<char *>(host) + ((size_t)i * hrows + (size_t)(k_lo - hk0)) * row;
        if (upload)
            LBM_CUDA(cudaMemcpy2DAsync(d, (size_t)L->pitch * e, h, row, row, (size_t)(k_hi - k_lo), cudaMemcpyHostToDevice, st));
        else
            LBM_CUDA(cudaMemcpy2DAsync(h, row, d, (size_t)L->pitch * e, row, (size_t)(k_hi - k_lo), cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

int host_step_resources(lb_lattice *L)
{
    if (!L->s_h2d) {
        LBM_CUDA(cudaStreamCreateWithFlags(&L->s_h2d, cudaStreamNonBlocking));
        LBM_CUDA(cudaStreamCreateWithFlags(&L->s_d2h, cudaStreamNonBlocking));
        LBM_CUDA(cudaEventCreateWithFlags(&L->ev_tail, cudaEventDisableTiming));
        LBM_CUDA(cudaEventCreateWithFlags(&L->ev_fin, cudaEventDisableTiming));
    }
    return 0;
}

// columns l = 0 and l = lny-1 of all nine populations, packed as cols[(i*2 + side)*lnx + k], into buffer `par`
template <typename T>
__global__ void scatter_cols_kernel(const __grid_constant__ StepParams<T> p, const T *__restrict__ cols, int par)
{
    T *buf = p.buf[par];
    const long long n = 18ll * p.lnx;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % p.lnx), side = (int)((t / p.lnx) & 1), i = (int)(t / (2ll * p.lnx));
        buf[(long long)i * p.pop_stride + (long long)(k + 1) * p.pitch + PAD_L + (side ? p.lny - 1 : 0)] = cols[t];
    }
}

// Phase 1 of a DECOMPOSED host step: the block's rim (rows 0 and lnx-1, columns 0 and lny-1) of the host input goes
// to the device first, so that lb_halo_refresh can push it into the neighbours' ghosts before any slab is computed.
template <typename T>
int step_host_begin(lb_lattice *L, const void *host_in)
{
    if (int r = host_step_resources(L)) return r;
    const int64_t lnx = L->cfg.lnx, lny = L->cfg.lny;
    const size_t col_bytes = (size_t)18 * lnx * sizeof(T);
    if (!L->h_cols) {
        LBM_CUDA(cudaMallocHost(&L->h_cols, col_bytes));
        LBM_CUDA(cudaMalloc(&L->d_cols, col_bytes));
    }
    void *hin = const_cast<void *>(host_in);
    const int par = L->cur;
    LBM_CUDA(cudaEventRecord(L->ev_fin, L->stream));
    LBM_CUDA(cudaStreamWaitEvent(L->s_h2d, L->ev_fin, 0));
    if (int r = copy_rows(L, hin, par, 0, 1, true, L->s_h2d)) return r;
    if (lnx > 1)
        if (int r = copy_rows(L, hin, par, lnx - 1, lnx, true, L->s_h2d)) return r;
    const T *h = static_cast<const T *>(host_in);
    T *pack = static_cast<T *>(L->h_cols);
    for (int i = 0; i < 9; ++i)
        for (int side = 0; side < 2; ++side) {
            const T *col = h + (size_t)i * lnx * lny + (side ? lny - 1 : 0);
            T *dst = pack + ((size_t)i * 2 + side) * lnx;
            for (int64_t k = 0; k < lnx; ++k) dst[k] = col[(size_t)k * lny];
        }
    LBM_CUDA(cudaMemcpyAsync(L->d_cols, L->h_cols, col_bytes, cudaMemcpyHostToDevice, L->s_h2d));
    scatter_cols_kernel<T><<<grid_for(18 * lnx, 256), 256, 0, L->s_h2d>>>(make_params<T>(L), static_cast<const T *>(L->d_cols), par);
    LBM_CUDA(cudaGetLastError());
    LBM_CUDA(cudaStreamSynchronize(L->s_h2d));
    L->launches++;
    L->host_begun = true;
    return 0;
}

// self_ring: the block closes its rings on itself and refreshes its own ghosts slab by slab; otherwise the ghosts
// of the current buffer were filled beforehand (step_host_begin + lb_halo_refresh on every rank + barriers).
template <typename T>
int step_host_pipelined(lb_lattice *L, const void *host_in, void *host_out, int nslabs, bool self_ring)
{
    const int64_t lnx = L->cfg.lnx;
    if (nslabs < 1) nslabs = 1;
    if (nslabs > lnx) nslabs = (int)lnx;
    if (int r = host_step_resources(L)) return r;
    while ((int)L->ev_up.size() < nslabs) {
        cudaEvent_t a, b;
        LBM_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        LBM_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        L->ev_up.push_back(a);
        L->ev_done.push_back(b);
    }
    const StepParams<T> p = make_params<T>(L);
    const int par = L->cur;
    void *hin = const_cast<void *>(host_in);
    auto lo = [&](int j) { return lnx * j / nslabs; };
    // everything already queued on the lattice's stream (previous steps) must precede the uploads
    LBM_CUDA(cudaEventRecord(L->ev_fin, L->stream));
    LBM_CUDA(cudaStreamWaitEvent(L->s_h2d, L->ev_fin, 0));
    LBM_CUDA(cudaStreamWaitEvent(L->s_d2h, L->ev_fin, 0));
    // the last row first: slab 0 pulls its periodic image (ghost row -1)
    if (self_ring)
        if (int r = copy_rows(L, hin, par, lnx - 1, lnx, true, L->s_h2d)) return r;
    LBM_CUDA(cudaEventRecord(L->ev_tail, L->s_h2d));
    for (int j = 0; j < nslabs; ++j) {
        if (int r = copy_rows(L, hin, par, lo(j), lo(j + 1), true, L->s_h2d)) return r;
        LBM_CUDA(cudaEventRecord(L->ev_up[j], L->s_h2d));
    }
    const long long n_rim = 2 * (L->cfg.lnx + L->cfg.lny);
    LBM_CUDA(cudaStreamWaitEvent(L->stream, L->ev_tail, 0));
    if (self_ring) halo_refresh_kernel<T><<<grid_for(n_rim, 256), 256, 0, L->stream>>>(p, (int)lnx - 1, (int)lnx);
    LBM_CUDA(cudaStreamWaitEvent(L->stream, L->ev_up[0], 0));
    if (self_ring) halo_refresh_kernel<T><<<grid_for(n_rim, 256), 256, 0, L->stream>>>(p, (int)lo(0), (int)lo(1));
    for (int j = 0; j < nslabs; ++j) {
        if (j + 1 < nslabs) {   // slab j pulls from the first row of slab j+1 (and its ghost columns)
            LBM_CUDA(cudaStreamWaitEvent(L->stream, L->ev_up[j + 1], 0));
            if (self_ring) halo_refresh_kernel<T><<<grid_for(n_rim, 256), 256, 0, L->stream>>>(p, (int)lo(j + 1), (int)lo(j + 2));
        }
        if (int r = launch_rows<T>(L, p, (int)lo(j), (int)lo(j + 1))) return r;
        LBM_CUDA(cudaEventRecord(L->ev_done[j], L->stream));
        LBM_CUDA(cudaStreamWaitEvent(L->s_d2h, L->ev_done[j], 0));
        if (int r = copy_rows(L, host_out, par ^ 1, lo(j), lo(j + 1), false, L->s_d2h)) return r;
        L->launches += 2;
    }
    advance_step_kernel<<<1, 1, 0, L->stream>>>(dev_state(L));
    L->launches += 2;
    L->steps++;
    L->cur ^= 1;
    LBM_CUDA(cudaGetLastError());
    LBM_CUDA(cudaEventRecord(L->ev_fin, L->s_d2h));
    LBM_CUDA(cudaStreamWaitEvent(L->stream, L->ev_fin, 0));
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    return 0;
}

}  // namespace

extern "C" {

/* One time step with HOST input and output (the reference's stateless calling convention:
 * cavity_opt2.py:275-277 on a host f_ikl): rows are uploaded, updated and downloaded slab by slab
 * on three streams, so H2D, compute and D2H overlap (PCIe full duplex).  host_out may alias host_in.
 * Single block with its ring closed on itself; pinned host memory is needed for real overlap. */
int lb_step_host(lb_lattice *L, const void *host_in, void *host_out, int nslabs)
{
    if (int r = check_ready(L)) return r;
    if (!host_in || !host_out) return lbm_fail(LB_ERR_INVALID, "null host buffer");
    if (L->inplace) return lbm_fail(LB_ERR_STATE, "lb_step_host is not available on in-place lattices");
    bool self_ring = true;
    for (int d = 0; d < LB_NUM_DIRS; ++d)
        if (L->nbr[d].base != L->base) self_ring = false;
    if (!self_ring && !L->host_begun)
        return lbm_fail(LB_ERR_STATE, "a decomposed host step needs lb_step_host_begin + lb_halo_refresh (and rank barriers) before lb_step_host");
    L->host_begun = false;
    LBM_ON_DEVICE(L);
    return L->cfg.dtype == LB_F64 ? step_host_pipelined<double>(L, host_in, host_out, nslabs, self_ring)
                                  : step_host_pipelined<float>(L, host_in, host_out, nslabs, self_ring);
}

/* First phase of a host step on a DECOMPOSED lattice (see lbm_b200.h). */
int lb_step_host_begin(lb_lattice *L, const void *host_in)
{
    if (int r = check_ready(L)) return r;
    if (!host_in) return lbm_fail(LB_ERR_INVALID, "null host buffer");
    if (L->cfg.boundary > LB_CAVITY_XPERIODIC) return lbm_fail(LB_ERR_INVALID, "lb_step_host supports the periodic and cavity boundaries");
    LBM_ON_DEVICE(L);
    return L->cfg.dtype == LB_F64 ? step_host_begin<double>(L, host_in) : step_host_begin<float>(L, host_in);
}

int lb_step_timed(lb_lattice *L, int64_t nsteps, float *ms)
{
    if (!L || !ms) return lbm_fail(LB_ERR_INVALID, "null argument");
    LBM_ON_DEVICE(L);
    LBM_CUDA(cudaEventRecord(L->ev0, L->stream));
    if (int r = lb_step(L, nsteps)) return r;
    LBM_CUDA(cudaEventRecord(L->ev1, L->stream));
    LBM_CUDA(cudaEventSynchronize(L->ev1));
    LBM_CUDA(cudaEventElapsedTime(ms, L->ev0, L->ev1));
    return 0;
}

int64_t lb_steps_done(lb_lattice *L) { return L ? L->steps : -1; }

int lb_health(lb_lattice *L)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    LBM_ON_DEVICE(L);
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    DevState h;
    LBM_CUDA(cudaMemcpy(&h, dev_state(L), sizeof(h), cudaMemcpyDeviceToHost));
    if (h.error) return lbm_fail(LB_ERR_HALO_TIMEOUT, "halo flag wait timed out inside the step kernel");
    if ((int64_t)h.step != L->steps)
        return lbm_fail(LB_ERR_STATE, "device step counter %llu != host %lld", h.step, (long long)L->steps);
    return 0;
}

int lb_moments(lb_lattice *L, void *rho, void *ux, void *uy)
{
    if (!L) return lbm_fail(LB_ERR_INVALID, "null lattice");
    LBM_ON_DEVICE(L);
    const long long n = L->cfg.lnx * L->cfg.lny;
    const size_t bytes = (size_t)n * L->elem;
    if (!L->d_mom) LBM_CUDA(cudaMalloc(&L->d_mom, 3 * bytes));
    char *d = static_cast<char *>(L->d_mom);
    if (L->cfg.boundary >= LB_SF_COUETTE)
        moments_kernel<double, true><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<double>(L), (double *)d, (double *)(d + bytes), (double *)(d + 2 * bytes));
    else if (L->cfg.dtype == LB_F64)
        moments_kernel<double, false><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<double>(L), (double *)d, (double *)(d + bytes), (double *)(d + 2 * bytes));
    else
        moments_kernel<float, false><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<float>(L), (float *)d, (float *)(d + bytes), (float *)(d + 2 * bytes));
    LBM_CUDA(cudaGetLastError());
    L->launches++;
    void *h[3] = {rho, ux, uy};
    for (int j = 0; j < 3; ++j)
        if (h[j]) LBM_CUDA(cudaMemcpyAsync(h[j], d + j * bytes, bytes, cudaMemcpyDeviceToHost, L->stream));
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    return 0;
}

int lb_probe_shear_enable(lb_lattice *L, int64_t l_global, const void *uy_k, int64_t capacity)
{
    if (!L || !uy_k || capacity < 1) return lbm_fail(LB_ERR_INVALID, "bad argument");
    const int64_t l_local = l_global - L->cfg.y0;
    if (l_local < 0 || l_local >= L->cfg.lny) return lbm_fail(LB_ERR_INVALID, "probe row is not inside this block");
    // The probe runs after every single step.  The stepping mode is a COLLECTIVE property of a decomposition, so a
    // block that was told to advance two steps per pass is not silently downgraded (its neighbours would wait for
    // frame flags it never posts): the caller switches every block to single steps first.
    if (L->temporal == 2)
        return lbm_fail(LB_ERR_STATE, "the shear probe needs single-step mode: lb_set_temporal(lat, 1, 0) on every block of the decomposition first");
    if (L->inplace) return lbm_fail(LB_ERR_STATE, "the shear probe is not available on in-place lattices");
    LBM_ON_DEVICE(L);
    if (L->d_uyk) cudaFree(L->d_uyk);
    if (L->d_series) cudaFree(L->d_series);
    if (L->d_prod) cudaFree(L->d_prod);
    L->d_uyk = L->d_series = L->d_prod = nullptr;
    LBM_CUDA(cudaMalloc(&L->d_prod, (size_t)4 * L->cfg.lnx * L->elem));      // two passes x two levels (resident kernels)
    LBM_CUDA(cudaMalloc(&L->d_uyk, (size_t)L->cfg.lnx * L->elem));
    LBM_CUDA(cudaMalloc(&L->d_series, (size_t)capacity * L->elem));
    LBM_CUDA(cudaMemcpy(L->d_uyk, uy_k, (size_t)L->cfg.lnx * L->elem, cudaMemcpyHostToDevice));
    LBM_CUDA(cudaMemset(L->d_series, 0, (size_t)capacity * L->elem));
    L->probe_capacity = capacity;
    L->probe_l_local = l_local;
    L->probe_step0 = L->steps;
    return 0;
}

int lb_probe_shear_read(lb_lattice *L, void *out, int64_t n)
{
    if (!L || !out || !L->d_series || n > L->probe_capacity) return lbm_fail(LB_ERR_INVALID, "bad argument");
    LBM_ON_DEVICE(L);
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    LBM_CUDA(cudaMemcpy(out, L->d_series, (size_t)n * L->elem, cudaMemcpyDeviceToHost));
    return 0;
}

/* Rows [k_lo, k_hi) of the current state to / from a C-contiguous host array (9, k_hi - k_lo, lny). */
static int rows_io(lb_lattice *L, int64_t k_lo, int64_t k_hi, void *host, bool upload)
{
    if (!L || !host) return lbm_fail(LB_ERR_INVALID, "null argument");
    if (k_lo < 0 || k_hi > L->cfg.lnx || k_lo >= k_hi) return lbm_fail(LB_ERR_INVALID, "row range [%lld, %lld) outside the block", (long long)k_lo, (long long)k_hi);
    LBM_ON_DEVICE(L);
    if (L->inplace && L->aa_swapped) {
        if (upload) return lbm_fail(LB_ERR_STATE, "partial uploads need the natural layout (an even number of steps since the last upload)");
        return aa_download_rows(L, host, k_lo, k_hi);
    }
    if (int r = copy_rows(L, host, L->cur, k_lo, k_hi, upload, L->stream, k_lo, k_hi - k_lo)) return r;
    LBM_CUDA(cudaStreamSynchronize(L->stream));
    return 0;
}
int lb_download_rows(lb_lattice *L, int64_t k_lo, int64_t k_hi, void *host) { return rows_io(L, k_lo, k_hi, host, false); }
int lb_upload_rows(lb_lattice *L, int64_t k_lo, int64_t k_hi, const void *host) { return rows_io(L, k_lo, k_hi, const_cast<void *>(host), true); }

int lb_checksum(lb_lattice *L, uint64_t *out)
{
    if (!L || !out) return lbm_fail(LB_ERR_INVALID, "null argument");
    LBM_ON_DEVICE(L);
    unsigned long long *d = nullptr;
    LBM_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long), L->stream);
    if (e == cudaSuccess) {
        const long long n = L->cfg.lnx * L->cfg.lny;
        if (L->cfg.dtype == LB_F64)
            checksum_kernel<double><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<double>(L), d);
        else
            checksum_kernel<float><<<grid_for(n, 256), 256, 0, L->stream>>>(make_params<float>(L), d);
        e = cudaGetLastError();
    }
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, L->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L->stream);
    cudaFree(d);
    LBM_CUDA(e);
    L->launches++;
    *out = (uint64_t)h;
    return 0;
}

int64_t lb_pitch(lb_lattice *L) { return L ? L->pitch : -1; }
int64_t lb_pop_stride(lb_lattice *L) { return L ? L->pop_stride : -1; }
int lb_kernel_launches(lb_lattice *L, int64_t *count)
{
    if (!L || !count) return lbm_fail(LB_ERR_INVALID, "null argument");
    *count = L->launches;
    return 0;
}

}  // extern "C"
