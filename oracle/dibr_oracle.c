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
