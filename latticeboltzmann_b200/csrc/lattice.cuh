// Device-side data model of one lattice block (shared by the kernels and the C ABI).
//
// HBM layout.  One cudaMalloc per block:
//
//     [ buffer A | buffer B | DevState ]
//
// Each buffer is a structure of arrays f[9][lnx + 2][pitch] in the reference's
// index order (channel, x = k, y = l; y fastest -- cavity_opt2.py:79-83):
//
//     element(i, k, l) = i * pop_stride + (k + 1) * pitch + (l + PAD_L)
//
// with k in [-1, lnx]: one ghost ROW on each x side, ALWAYS present (a single block
// closes the periodic ring on itself).  PAD_L puts cell l = 0 on a 128-byte boundary
// and `pitch` is a multiple of 128 bytes, so every row of every population starts
// line-aligned and all stores of the pull scheme are aligned.
//
// The ghost COLUMNS (l = -1 and l = lny) are NOT stored in the rows: they live in small
// contiguous side arrays ("ycol"), one per buffer,
//
//     ycol(side, j, k) = side*3*(lnx+2) + j*(lnx+2) + (k+1),  k in [-1, lnx]
//     side 0 = column l = -1  holding j = 0,1,2 -> N, NE, NW  (pulled by the cells of l = 0)
//     side 1 = column l = lny holding j = 0,1,2 -> S, SW, SE  (pulled by the cells of l = lny-1)
//
// so that a y-neighbour's halo push is a COALESCED store (consecutive k -> consecutive
// addresses) instead of one 8-byte NVLink transaction per row (measured: strided peer
// stores cost 1.34 ms per step at 16384 rows; contiguous ones are free).
#pragma once
#include <stdint.h>

namespace lbm {

constexpr int PAD_L = 32;          // elements before l = 0 in a row (128 B for f32, 256 B for f64)
constexpr int TILE_L = 256;        // threads per CTA = cells of one row handled by a CTA
constexpr int NUM_DIRS = 8;

// (dx, dy) of direction slot d (include/lbm_b200.h).
__host__ __device__ constexpr int dir_dx(int d) { return d == 0 || d == 4 || d == 5 ? -1 : (d == 1 || d == 6 || d == 7 ? 1 : 0); }
__host__ __device__ constexpr int dir_dy(int d) { return d == 2 || d == 4 || d == 6 ? -1 : (d == 3 || d == 5 || d == 7 ? 1 : 0); }
__host__ __device__ constexpr int dir_opp(int d) { return d == 0 ? 1 : d == 1 ? 0 : d == 2 ? 3 : d == 3 ? 2 : d == 4 ? 7 : d == 7 ? 4 : d == 5 ? 6 : 5; }

// Lives at the tail of the block's allocation so that neighbours (other
// processes / GPUs) can post their "halo pushed" flags straight into it.
struct DevState {
    unsigned long long step;               // completed time steps
    unsigned int cur;                      // which buffer (0 = A, 1 = B) holds the current state; flips once per PASS
    unsigned int pad0;
    unsigned long long flag_in[NUM_DIRS];  // flag_in[d]: steps whose halos the neighbour in slot d has pushed
    unsigned int edge_done;                // edge CTAs finished in the running launch
    unsigned int all_done;                 // CTAs finished in the running launch
    unsigned int error;                    // != 0: a halo flag wait timed out
    unsigned int frame_done;               // temporal blocking: frame-kernel CTAs finished in the running launch
    unsigned long long fflag_in[NUM_DIRS]; // temporal blocking: level-(n+1) frame ghosts pushed by the neighbour in slot d
    unsigned long long grid_bar;           // resident multi-step kernel: arrivals at its grid barrier (monotone)
    unsigned int t2_done;                  // temporal blocking: fused-interior CTAs finished in the running launch
    unsigned int pass_done;                // temporal blocking: parts of the running pass (interior, frame level n+2) that are complete
};
static_assert(sizeof(DevState) <= 256, "the allocation reserves 256 bytes for DevState");

// ---- temporal blocking (two time steps per pass over HBM), temporal.cuh ----------------------------
// The FRAME of a block = its cells closer than 3 cells to the perimeter.  Its level-(n+1) values live in a
// small side storage next to the buffers (elements, per block):
//     top    [9][3][pitch]      rows k = 0..2            element (i,k,l) -> (i*3 + k)*pitch + l + PAD_L
//     bottom [9][3][pitch]      rows k = lnx-3..lnx-1
//     left   [9][lnx][4]        columns l = 0..2         element (i,k,l) -> (i*lnx + k)*4 + l
//     right  [9][lnx][4]        columns l = lny-3..lny-1
//     grow   [2][3][pitch]      level-(n+1) ghost rows k = -1 (E,NE,SE) and k = lnx (W,NW,SW)
//     gcol   [2][3][lnx+2]      level-(n+1) ghost columns, same convention as ycol
constexpr int FRAME_W = 3;
__host__ __device__ constexpr int xrow_slot(int i) { return (i == 1 || i == 3) ? 0 : ((i == 5 || i == 6) ? 1 : 2); }
template <typename T>
struct FrameView {
    T *top, *bottom, *left, *right, *grow, *gcol;
};
__host__ __device__ inline long long frame_elems(long long lnx, long long pitch)
{
    return 2 * 27 * pitch + 2 * 36 * lnx + 6 * pitch + 6 * (lnx + 2);
}
template <typename T>
__host__ __device__ inline FrameView<T> frame_view(T *base, long long lnx, long long pitch)
{
    FrameView<T> f;
    f.top = base;
    f.bottom = f.top + 27 * pitch;
    f.left = f.bottom + 27 * pitch;
    f.right = f.left + 36 * lnx;
    f.grow = f.right + 36 * lnx;
    f.gcol = f.grow + 6 * pitch;
    return f;
}

__host__ __device__ constexpr int ycol_slot(int i) { return (i == 2 || i == 4) ? 0 : ((i == 5 || i == 7) ? 1 : 2); }

template <typename T>
struct NbrView {
    T *buf[2];                       // neighbour's buffers A / B (local, peer or IPC-mapped address)
    T *ycol[2];                      // neighbour's ghost-column arrays for buffers A / B
    unsigned long long *flag_in;     // neighbour's DevState::flag_in
    unsigned long long *fflag_in;    // neighbour's DevState::fflag_in
    T *frame;                        // neighbour's frame storage (temporal blocking)
    long long pop_stride, pitch;
    int lnx, lny;
};

template <typename T>
struct StepParams {
    T *buf[2];
    T *ycol[2];
    T *frame;                // frame storage (temporal blocking)
    DevState *st;
    long long pop_stride, pitch;
    long long x0, y0, gnx, gny;
    int lnx, lny;
    int tiles_l, tiles_k, rows_per_tile;
    long long n_perimeter;   // cells on the block's perimeter
    int n_rim_ctas;          // CTAs [0, n_rim_ctas) handle the perimeter
    T omega, u_wall;
    T sf_uw6;                // simple_flows: (1/6)*uw, evaluated in double on the host (PoiseuilleFlow.py:73-74)
    T rho_in, rho_out;       // simple_flows Poiseuille (PoiseuilleFlow.py:134-135)
    int t2_rows, t2_tiles_l, t2_tiles_k;   // temporal blocking: rows per fused tile, tile grid over the deep interior
    int all_rim;             // 1: every cell takes the general (rim) path, no interior tiles
    int bc;                  // boundary kind (BoundaryKind) for code that is not templated on it
    int aa_swapped;          // in-place lattices: 1 while the single buffer is in the swapped layout (aa.cuh)
    int sys_scope;           // 1: some neighbour is on another device / process (system-scope fences)
    unsigned long long halo_timeout_ns;   // give up waiting for a neighbour's flag after this long
    // Byte offsets relative to a cell's own slot, precomputed on the host so that the kernel adds
    // them straight from the constant bank (no registers): ld_off[i] addresses the pull source
    // (i, k - cx_i, l - cy_i).
    long long ld_off[9];
    // per-cell boundary table (BC_SF_TABLE): tab_n cells (flat k*lny + l), for each the element offsets of its nine
    // pre-stream sources inside a buffer and nine additive constants; one bit per lattice cell marks the listed ones
    const int *tab_cells;
    const long long *tab_src;
    const T *tab_add;
    const unsigned int *tab_mask;
    const int *tab_rank;     // listed cells before each 32-cell word of tab_mask: table index of cell c = tab_rank[c >> 5] + popc(lower bits)
    int tab_n;
    NbrView<T> nbr[NUM_DIRS];
};

}  // namespace lbm
