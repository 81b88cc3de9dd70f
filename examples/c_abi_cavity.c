/*
 * Minimal C client of the drop-in boundary (include/lbm_b200.h): a lid-driven cavity advanced on the GPU
 * through the plain C ABI, no Python, no torch.  This is the call sequence a native host (C, C++, Fortran,
 * cgo, JNI ...) uses; it mirrors the main loop of simulators/parallel_lid_drive_cavity/cavity_opt2.py:265-283.
 *
 *   gcc -O2 -Iinclude examples/c_abi_cavity.c -o c_abi_cavity \
 *       -Llatticeboltzmann_b200/csrc -llbm_b200 -Wl,-rpath,$PWD/latticeboltzmann_b200/csrc
 *   ./c_abi_cavity [nx ny nsteps]
 *
 * Exit code 3 = no CUDA device (the library has no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lbm_b200.h"

#define CHECK(call)                                                        \
    do {                                                                   \
        int rc_ = (call);                                                  \
        if (rc_ != LB_OK) {                                                \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, lb_last_error()); \
            return rc_ == LB_ERR_NO_DEVICE ? 3 : 1;                        \
        }                                                                  \
    } while (0)

int main(int argc, char **argv)
{
    const int64_t nx = argc > 1 ? atoll(argv[1]) : 256, ny = argc > 2 ? atoll(argv[2]) : 256;
    const int64_t nsteps = argc > 3 ? atoll(argv[3]) : 1000;

    lb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0;
    cfg.dtype = LB_F64;
    cfg.boundary = LB_CAVITY;            /* four bounce-back walls + moving lid, cavity_opt2.py:109-177 */
    cfg.arith = LB_ARITH_EXACT;          /* bit-identical to the reference's no-FMA build                */
    cfg.gnx = cfg.lnx = nx;              /* one block = the whole lattice                                */
    cfg.gny = cfg.lny = ny;
    cfg.omega = 1.7;                     /* cavity_opt2.py:66                                            */
    cfg.u_wall = 0.1;                    /* cavity_opt2.py:109                                           */

    lb_lattice *lat = NULL;
    CHECK(lb_create(&cfg, &lat));
    lb_export self;
    CHECK(lb_get_export(lat, &self));
    for (int d = 0; d < LB_NUM_DIRS; ++d) CHECK(lb_connect(lat, d, &self));   /* periodic ring closed on itself */
    CHECK(lb_init_equilibrium(lat, NULL, NULL, NULL));                          /* rho = 1, u = 0, :265-269        */
    CHECK(lb_halo_refresh(lat));

    float ms = 0.f;
    CHECK(lb_step_timed(lat, nsteps, &ms));                                     /* :272-277, fused on the device   */
    CHECK(lb_health(lat));

    double *rho = malloc(sizeof(double) * nx * ny), *ux = malloc(sizeof(double) * nx * ny);
    CHECK(lb_moments(lat, rho, ux, NULL));                                      /* :280-281                        */
    double mass = 0, umax = 0;
    for (int64_t i = 0; i < nx * ny; ++i) {
        mass += rho[i];
        if (ux[i] > umax) umax = ux[i];
    }
    uint64_t digest = 0;
    CHECK(lb_checksum(lat, &digest));                                           /* 64-bit digest of f, computed on the device */
    printf("%lld x %lld cavity, %lld steps: %.3f ms, %.1f MLUPS, mass %.6f, max ux %.6f, digest %016llx\n", (long long)nx, (long long)ny,
           (long long)nsteps, ms, nx * ny * (double)nsteps / (ms * 1e-3) / 1e6, mass, umax, (unsigned long long)digest);
    free(rho);
    free(ux);
    CHECK(lb_destroy(lat));

    /* The same run with ONE copy of the populations (in-place AA pattern): same bits, half the memory. */
    uint64_t digest_inplace = 0;
    CHECK(lb_create_ex(&cfg, LB_CREATE_INPLACE, &lat));
    CHECK(lb_get_export(lat, &self));
    for (int d = 0; d < LB_NUM_DIRS; ++d) CHECK(lb_connect(lat, d, &self));
    CHECK(lb_init_equilibrium(lat, NULL, NULL, NULL));
    CHECK(lb_step(lat, nsteps));
    CHECK(lb_health(lat));
    CHECK(lb_checksum(lat, &digest_inplace));
    CHECK(lb_destroy(lat));
    printf("in-place lattice: digest %016llx (%s)\n", (unsigned long long)digest_inplace, digest_inplace == digest ? "identical" : "DIFFERENT");
    return digest_inplace == digest ? 0 : 2;
}
