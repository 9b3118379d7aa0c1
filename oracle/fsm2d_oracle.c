/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, flat-array CPU restatement of the reference's 2-D rectilinear fast-sweeping solver Grid2Drnfs (node
 * slowness; all five node updates: square cells, rotated template, dx != dz, WENO, WENO with dx != dz), file:line
 * citations in fsm2d_oracle_impl.h.  SURVEY section 8 row f4 (the 2-D twins) is NOT built in the product yet: this is
 * the oracle that row will be tested against, pinned bit for bit against the unmodified reference
 * (tests/test_oracle.py::test_2d_*, live where oracle/_ref exists).  Only tests/ may load it.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

#define REAL double
#define SFX _d
#define REAL_MAX DBL_MAX
#define REAL_EPS DBL_EPSILON
#include "fsm2d_oracle_impl.h"
#undef REAL
#undef SFX
#undef REAL_MAX
#undef REAL_EPS

#define REAL float
#define SFX _f
#define REAL_MAX FLT_MAX
#define REAL_EPS FLT_EPSILON
#include "fsm2d_oracle_impl.h"
#undef REAL
#undef SFX
#undef REAL_MAX
#undef REAL_EPS
