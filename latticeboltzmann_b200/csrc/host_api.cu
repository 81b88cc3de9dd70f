// C ABI group (1): the reference's stateless kernels on HOST buffers
// (c/_lbkernels.cpp:36-48 and PyLB/Streaming.py:33-46), executed on the GPU.
//
// Contract mirrored from the reference: results are written in
// place into the caller's buffers; there is no length check between rho/ux/uy
// and f beyond `n`; rho = 0 silently produces inf/NaN.
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/lbm_b200.h"
#include "api_common.h"
#include "d2q9_math.cuh"

using namespace lbm;

namespace {

template <typename T>
__global__ void k_equilibriumn(const T *__restrict__ rho, const T *__restrict__ ux, const T *__restrict__ uy,
                               T *__restrict__ f, long long n)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        T e[9];
        d2q9_equilibrium<T, true>(rho[t], ux[t], uy[t], e);
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i * n + t] = e[i];
    }
}

template <typename T>
__global__ void k_collide(T *__restrict__ f, long long n, T omega)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        T v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = f[i * n + t];
        d2q9_collide<T, true>(v, omega);
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i * n + t] = v[i];
    }
}

// out[i,k,l] = in[i,(k-cx) mod nx,(l-cy) mod ny]  (np.roll, PyLB/Streaming.py:45-46)
template <typename T>
__global__ void k_stream(const T *__restrict__ in, T *__restrict__ out, long long nx, long long ny)
{
    const long long n = nx * ny;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long k = t / ny, l = t % ny;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            long long ks = k - cx_of(i), ls = l - cy_of(i);
            ks = ks < 0 ? ks + nx : (ks >= nx ? ks - nx : ks);
            ls = ls < 0 ? ls + ny : (ls >= ny ? ls - ny : ls);
            out[i * n + t] = in[i * n + ks * ny + ls];
        }
    }
}

// Self-check of rn_div_const against the IEEE division intrinsic: counts mismatching bit patterns.
template <typename T, typename U>
__global__ void k_check_div_const(const U *bits, long long n, unsigned long long *bad)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        T x;
        U b = bits[t];
        memcpy(&x, &b, sizeof(T));
        T a9 = rn_div_const<9>(x), a6 = rn_div_const<6>(x);
        T r9 = sizeof(T) == 8 ? (T)__ddiv_rn((double)x, 9.0) : (T)__fdiv_rn((float)x, 9.0f);
        T r6 = sizeof(T) == 8 ? (T)__ddiv_rn((double)x, 6.0) : (T)__fdiv_rn((float)x, 6.0f);
        U ua9, ua6, ur9, ur6;
        memcpy(&ua9, &a9, sizeof(T)); memcpy(&ua6, &a6, sizeof(T)); memcpy(&ur9, &r9, sizeof(T)); memcpy(&ur6, &r6, sizeof(T));
        const bool nan9 = (a9 != a9) && (r9 != r9), nan6 = (a6 != a6) && (r6 != r6);
        if ((ua9 != ur9 && !nan9) || (ua6 != ur6 && !nan6)) atomicAdd(bad, 1ull);
    }
}

int grid_for(long long n) { long long g = (n + 255) / 256; return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g)); }

int need_device()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return lbm_fail(LB_ERR_NO_DEVICE, "no CUDA device visible: liblbm_b200 has no CPU fallback");
    }
    return 0;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

template <typename T>
int equilibrium1(T rho, T ux, T uy, T *out9)
{
    if (!out9) return lbm_fail(LB_ERR_INVALID, "null argument");
    if (int r = need_device()) return r;
    DevBuf in, out;
    LBM_CUDA(cudaMalloc(&in.p, 3 * sizeof(T)));
    LBM_CUDA(cudaMalloc(&out.p, 9 * sizeof(T)));
    const T h[3] = {rho, ux, uy};
    LBM_CUDA(cudaMemcpy(in.p, h, sizeof(h), cudaMemcpyHostToDevice));
    T *d = static_cast<T *>(in.p);
    k_equilibriumn<T><<<1, 32>>>(d, d + 1, d + 2, static_cast<T *>(out.p), 1);
    LBM_CUDA(cudaGetLastError());
    LBM_CUDA(cudaMemcpy(out9, out.p, 9 * sizeof(T), cudaMemcpyDeviceToHost));
    return 0;
}

template <typename T>
int equilibriumn(const T *rho, const T *ux, const T *uy, T *f, int64_t n)
{
    if (n < 0 || (n > 0 && (!rho || !ux || !uy || !f))) return lbm_fail(LB_ERR_INVALID, "bad argument");
    if (n == 0) return 0;
    if (int r = need_device()) return r;
    DevBuf in, out;
    const size_t b = (size_t)n * sizeof(T);
    LBM_CUDA(cudaMalloc(&in.p, 3 * b));
    LBM_CUDA(cudaMalloc(&out.p, 9 * b));
    char *d = static_cast<char *>(in.p);
    LBM_CUDA(cudaMemcpy(d, rho, b, cudaMemcpyHostToDevice));
    LBM_CUDA(cudaMemcpy(d + b, ux, b, cudaMemcpyHostToDevice));
    LBM_CUDA(cudaMemcpy(d + 2 * b, uy, b, cudaMemcpyHostToDevice));
    k_equilibriumn<T><<<grid_for(n), 256>>>((const T *)d, (const T *)(d + b), (const T *)(d + 2 * b), static_cast<T *>(out.p), n);
    LBM_CUDA(cudaGetLastError());
    LBM_CUDA(cudaMemcpy(f, out.p, 9 * b, cudaMemcpyDeviceToHost));
    return 0;
}

template <typename T>
int collide(T *f, int64_t n, T omega)
{
    if (n < 0 || (n > 0 && !f)) return lbm_fail(LB_ERR_INVALID, "bad argument");
    if (n == 0) return 0;
    if (int r = need_device()) return r;
    DevBuf d;
    const size_t b = (size_t)9 * n * sizeof(T);
    LBM_CUDA(cudaMalloc(&d.p, b));
    LBM_CUDA(cudaMemcpy(d.p, f, b, cudaMemcpyHostToDevice));
    k_collide<T><<<grid_for(n), 256>>>(static_cast<T *>(d.p), n, omega);
    LBM_CUDA(cudaGetLastError());
    LBM_CUDA(cudaMemcpy(f, d.p, b, cudaMemcpyDeviceToHost));
    return 0;
}

template <typename T>
int stream(T *f, int64_t nx, int64_t ny)
{
    if (nx < 0 || ny < 0 || (nx * ny > 0 && !f)) return lbm_fail(LB_ERR_INVALID, "bad argument");
    if (nx * ny == 0) return 0;
    if (int r = need_device()) return r;
    DevBuf in, out;
    const size_t b = (size_t)9 * nx * ny * sizeof(T);
    LBM_CUDA(cudaMalloc(&in.p, b));
    LBM_CUDA(cudaMalloc(&out.p, b));
    LBM_CUDA(cudaMemcpy(in.p, f, b, cudaMemcpyHostToDevice));
    k_stream<T><<<grid_for(nx * ny), 256>>>(static_cast<const T *>(in.p), static_cast<T *>(out.p), nx, ny);
    LBM_CUDA(cudaGetLastError());
    LBM_CUDA(cudaMemcpy(f, out.p, b, cudaMemcpyDeviceToHost));
    return 0;
}

// The opt2 loop body on a host array, through the device-resident lattice.
int step_host(void *f, int dtype, int64_t nx, int64_t ny, int boundary, double omega, double u0, int64_t nsteps)
{
    if (!f || nx < 1 || ny < 1 || nsteps < 0) return lbm_fail(LB_ERR_INVALID, "bad argument");
    if (boundary != LB_PERIODIC && boundary != LB_CAVITY && boundary != LB_CAVITY_XPERIODIC)
        return lbm_fail(LB_ERR_INVALID, "lbk_step_host supports LB_PERIODIC, LB_CAVITY and LB_CAVITY_XPERIODIC");
    if (int r = need_device()) return r;
    lb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    int dev = 0;
    LBM_CUDA(cudaGetDevice(&dev));
    cfg.device = dev;
    cfg.dtype = dtype;
    cfg.boundary = boundary;
    cfg.arith = LB_ARITH_EXACT;
    cfg.gnx = cfg.lnx = nx;
    cfg.gny = cfg.lny = ny;
    cfg.omega = omega;
    cfg.u_wall = u0;
    lb_lattice *L = nullptr;
    int r = lb_create(&cfg, &L);
    if (r) return r;
    lb_export self;
    r = lb_get_export(L, &self);
    for (int d = 0; !r && d < LB_NUM_DIRS; ++d) r = lb_connect(L, d, &self);
    if (nsteps == 1) {
        if (!r) r = lb_step_host(L, f, f, 16);      // H2D / compute / D2H overlapped slab by slab
    } else {
        if (!r) r = lb_upload_f(L, f);
        if (!r) r = lb_halo_refresh(L);
        if (!r) r = lb_step(L, nsteps);
        if (!r) r = lb_download_f(L, f);
    }
    if (!r) r = lb_health(L);
    lb_destroy(L);
    return r;
}

}  // namespace

template <typename T, typename U>
int check_div_const(const U *bits, int64_t n, int64_t *mismatches)
{
    if (!bits || !mismatches || n < 0) return lbm_fail(LB_ERR_INVALID, "bad argument");
    if (int r = need_device()) return r;
    DevBuf d, b;
    LBM_CUDA(cudaMalloc(&d.p, (size_t)n * sizeof(U) + 8));
    LBM_CUDA(cudaMalloc(&b.p, 8));
    LBM_CUDA(cudaMemset(b.p, 0, 8));
    LBM_CUDA(cudaMemcpy(d.p, bits, (size_t)n * sizeof(U), cudaMemcpyHostToDevice));
    k_check_div_const<T, U><<<grid_for(n), 256>>>(static_cast<const U *>(d.p), n, static_cast<unsigned long long *>(b.p));
    LBM_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    LBM_CUDA(cudaMemcpy(&h, b.p, 8, cudaMemcpyDeviceToHost));
    *mismatches = (int64_t)h;
    return 0;
}

extern "C" {

int lbk_selftest_div_const_f64(const uint64_t *bits, int64_t n, int64_t *mismatches) { return check_div_const<double, unsigned long long>(reinterpret_cast<const unsigned long long *>(bits), n, mismatches); }
int lbk_selftest_div_const_f32(const uint32_t *bits, int64_t n, int64_t *mismatches) { return check_div_const<float, unsigned int>(bits, n, mismatches); }

int lbk_equilibrium1_f32(float rho, float ux, float uy, float *out9) { return equilibrium1<float>(rho, ux, uy, out9); }
int lbk_equilibrium1_f64(double rho, double ux, double uy, double *out9) { return equilibrium1<double>(rho, ux, uy, out9); }
int lbk_equilibriumn_f32(const float *rho, const float *ux, const float *uy, float *f, int64_t n) { return equilibriumn<float>(rho, ux, uy, f, n); }
int lbk_equilibriumn_f64(const double *rho, const double *ux, const double *uy, double *f, int64_t n) { return equilibriumn<double>(rho, ux, uy, f, n); }
int lbk_collide_f32(float *f, int64_t n, float omega) { return collide<float>(f, n, omega); }
int lbk_collide_f64(double *f, int64_t n, double omega) { return collide<double>(f, n, omega); }
int lbk_stream_f32(float *f, int64_t nx, int64_t ny) { return stream<float>(f, nx, ny); }
int lbk_stream_f64(double *f, int64_t nx, int64_t ny) { return stream<double>(f, nx, ny); }
int lbk_step_host_f32(float *f, int64_t nx, int64_t ny, int boundary, float omega, float u0, int64_t nsteps)
{
    return step_host(f, LB_F32, nx, ny, boundary, omega, u0, nsteps);
}
int lbk_step_host_f64(double *f, int64_t nx, int64_t ny, int boundary, double omega, double u0, int64_t nsteps)
{
    return step_host(f, LB_F64, nx, ny, boundary, omega, u0, nsteps);
}

}  // extern "C"
