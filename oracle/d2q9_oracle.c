/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the D2Q9 BGK time step
 * (plain-C restatement of the reference's opt2 path; see d2q9_oracle_impl.h
 * for the per-function reference citations and the parity status).
 *
 * Build:  make -C oracle     (gcc -O2 -ffp-contract=off, no -march)
 */
#include <stdint.h>
#include <string.h>

#define T float
#define SUF f32
#include "d2q9_oracle_impl.h"
#undef T
#undef SUF

#define T double
#define SUF f64
#include "d2q9_oracle_impl.h"
#undef T
#undef SUF

int orc_abi_version(void) { return 1; }
