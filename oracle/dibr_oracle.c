/*
 * ORACLE -- test infrastructure only.  Nothing under self6dpp_b200/ may import, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 * See dibr_oracle_body.h for what it restates and the PARITY UNPINNED statement.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: the compiler must not fuse or re-associate;
 * every FMA in the float build is written explicitly.)
 */
#include <math.h>

#define REAL float
#define SUFFIX _f32
#define FMA(a, b, c) fmaf((a), (b), (c))
#define EXP(x) expf(x)
#include "dibr_oracle_body.h"
#undef REAL
#undef SUFFIX
#undef FMA
#undef EXP

#define REAL double
#define SUFFIX _f64
#define FMA(a, b, c) fma((a), (b), (c))
#define EXP(x) exp(x)
#include "dibr_oracle_body.h"
#undef REAL
#undef SUFFIX
#undef FMA
#undef EXP

int dibr_oracle_abi_version(void) { return 1; }

/* OpenMP threads the pixel loops use.  bench.py sets this explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers,
 * which would silently turn the CPU arm into a single-thread run. */
#ifdef _OPENMP
#include <omp.h>
void dibr_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int dibr_oracle_get_threads(void) { return omp_get_max_threads(); }
#else
void dibr_oracle_set_threads(int n) { (void)n; }
int dibr_oracle_get_threads(void) { return 1; }
#endif
