/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, flat-array (SoA) CPU restatement of the reference's 3D rectilinear
 * fast-sweeping path: Grid3Drnfs / Grid3Drcfs drivers, Grid3Drn::sweep,
 * update_node, sweep_weno3, update_node_weno3, weno3_upwind, initFSM,
 * getTraveltime, getTraveltimeFromRaypath / grad / computeSlowness (file:line citations in
 * fsm_oracle_impl.h).
 *
 * It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (ttcr_b200/) never links, imports or calls anything in this directory.
 *
 * PARITY PINNED: tests/test_oracle.py checks this restatement bit-for-bit
 * (double and float, first-order and WENO, node and cell slowness, on-node and
 * off-node sources, lexicographic and plane order) against the unmodified
 * reference compiled from /root/reference into oracle/_ref/ (ref_shim.cpp), and
 * against the committed golden fields in tests/golden/ that the same reference
 * build produced (oracle/make_golden.py), and against the reference's own
 * acceptance criterion (mean relative error vs the analytic solution < 0.01 at
 * the rcv.dat points, tests/test_grid3d.cpp:68-96,181,199).
 *
 * Why it exists beside oracle/_ref: (1) the reference needs ~257 B/node, so
 * 1024^3 cannot run; this needs 2 arrays + 1 byte; (2) /root/reference does not
 * exist on the GPU box, this file does.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

#define REAL double
#define SFX _d
#define FABS fabs
#define SQRT sqrt
#define REAL_MAX DBL_MAX
#define REAL_EPS DBL_EPSILON
#include "fsm_oracle_impl.h"
#undef REAL
#undef SFX
#undef FABS
#undef SQRT
#undef REAL_MAX
#undef REAL_EPS

#define REAL float
#define SFX _f
#define FABS fabsf
#define SQRT sqrtf
#define REAL_MAX FLT_MAX
#define REAL_EPS FLT_EPSILON
#include "fsm_oracle_impl.h"
#undef REAL
#undef SFX
#undef FABS
#undef SQRT
#undef REAL_MAX
#undef REAL_EPS
