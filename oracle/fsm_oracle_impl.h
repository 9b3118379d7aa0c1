/* TEST INFRASTRUCTURE ONLY -- body of the plain-C restatement, included twice by
 * fsm_oracle.c (REAL = double, REAL = float).  See fsm_oracle.c for the header.
 *
 * All file:line citations are relative to /root/reference/ttcr/.
 *
 * Arithmetic types deliberately mirror the reference template instantiated at
 * T1 = REAL: variables are REAL, literals are `double`, so mixed expressions are
 * evaluated in double and rounded to REAL on assignment exactly as the C++
 * does (usual arithmetic conversions are the same in C and C++; compile with
 * -ffp-contract=off so that no FMA is formed, as in the reference x86-64 build).
 */

#ifndef REAL
#error "define REAL, SFX and REAL_MAX / REAL_EPS before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct {
    size_t nx1, ny1, nz1; /* node counts (ncx+1, ...) */
    REAL dx;
    REAL *tt;
    const REAL *s;
} FN(grid_);

#define NIDX(g, i, j, k) ((((size_t)(k)) * (g)->ny1 + (size_t)(j)) * (g)->nx1 + (size_t)(i))

/* ---- Grid3Drcfs::setSlowness, Grid3Drcfs.h:88-171 ------------------------
 * node slowness = mean of the 1/2/4/8 adjacent cells.  Summation order follows
 * the source: cells enumerated with k outer, j, i inner (":157-170"), except on
 * the x = const faces where the source enumerates j outer, k inner (":143-154").
 * The weights 1, .5, .25, .125 are powers of two, so only the order of the
 * additions can influence the last bit. */
void FN(fsmo_cell_to_node)(const REAL *sc, size_t ncx, size_t ncy, size_t ncz, REAL *sn) {
    const size_t nx1 = ncx + 1, ny1 = ncy + 1;
    for (size_t k = 0; k <= ncz; ++k)
        for (size_t j = 0; j <= ncy; ++j)
            for (size_t i = 0; i <= ncx; ++i) {
                long ks[2], js[2], is[2];
                int nk = 0, nj = 0, ni = 0;
                if (k < ncz) ks[nk++] = (long)k;
                if (k > 0) ks[nk++] = (long)k - 1;
                if (j < ncy) js[nj++] = (long)j;
                if (j > 0) js[nj++] = (long)j - 1;
                if (i < ncx) is[ni++] = (long)i;
                if (i > 0) is[ni++] = (long)i - 1;
                REAL sum = 0;
                int first = 1;
                if (ni == 1 && nj == 2 && nk == 2) {
                    /* x face: j outer, k inner (Grid3Drcfs.h:143-154) */
                    for (int b = 0; b < nj; ++b)
                        for (int a = 0; a < nk; ++a) {
                            REAL v = sc[((size_t)ks[a] * ncy + (size_t)js[b]) * ncx + (size_t)is[0]];
                            if (first) { sum = v; first = 0; } else sum = sum + v;
                        }
                } else {
                    for (int a = 0; a < nk; ++a)
                        for (int b = 0; b < nj; ++b)
                            for (int c = 0; c < ni; ++c) {
                                REAL v = sc[((size_t)ks[a] * ncy + (size_t)js[b]) * ncx + (size_t)is[c]];
                                if (first) { sum = v; first = 0; } else sum = sum + v;
                            }
                }
                const int n = nk * nj * ni;
                REAL out;
                if (n == 1) out = sum;
                else if (n == 2) out = 0.5 * sum;
                else if (n == 4) out = 0.25 * sum;
                else out = 0.125 * sum;
                sn[(k * ny1 + j) * nx1 + i] = out;
            }
}

/* ---- Godunov cascade shared by update_node and update_node_weno3 ---------
 * Grid3Drn.h:2936-2957 (and :3462-3483): sort3, 1D -> 2D -> 3D, keep if smaller */
static inline void FN(cascade_)(FN(grid_) * g, size_t n, REAL a1, REAL a2, REAL a3) {
    REAL t;
    if (a1 > a2) { t = a1; a1 = a2; a2 = t; }
    if (a1 > a3) { t = a1; a1 = a3; a3 = t; }
    if (a2 > a3) { t = a2; a2 = a3; a3 = t; }
    REAL fh = g->s[n] * g->dx;
    t = a1 + fh;
    if (t > a2) {
        t = 0.5 * (a1 + a2 + sqrt(2. * fh * fh - (a1 - a2) * (a1 - a2)));
        if (t > a3) {
            t = 1. / 3. * ((a1 + a2 + a3) + sqrt(-2. * a1 * a1 + 2. * a1 * a2 - 2. * a2 * a2 +
                                                 2. * a1 * a3 + 2. * a2 * a3 -
                                                 2. * a3 * a3 + 3. * fh * fh));
        }
    }
    if (t < g->tt[n]) g->tt[n] = t;
}

/* first-order one-sided minimum along one axis: Grid3Drn.h:2906-2934 */
static inline REAL FN(axis1_)(const REAL *tt, size_t n, size_t q, size_t nc, size_t stride) {
    if (q == 0) return tt[n + stride];
    if (q == nc) return tt[n - stride];
    REAL a = tt[n - stride];
    REAL t = tt[n + stride];
    return a < t ? a : t;
}

/* Grid3Drn.h:3047-3075 */
static inline REAL FN(weno3_)(REAL v0, REAL v1, REAL v2, REAL v3, REAL v4, REAL dx, int forward) {
    const REAL eps = REAL_EPS;
    if (forward) {
        const REAL num = (v4 - 2.0 * v3 + v2);
        const REAL den = (v3 - 2.0 * v2 + v1);
        const REAL r = (eps + num * num) / (eps + den * den);
        const REAL w = 1.0 / (1.0 + 2.0 * r * r);
        const REAL ap = (1.0 - w) * (v3 - v1) / (2.0 * dx) + w * (-v4 + 4.0 * v3 - 3.0 * v2) / (2.0 * dx);
        return v2 + dx * ap;
    } else {
        const REAL num = (v2 - 2.0 * v1 + v0);
        const REAL den = (v3 - 2.0 * v2 + v1);
        const REAL r = (eps + num * num) / (eps + den * den);
        const REAL w = 1.0 / (1.0 + 2.0 * r * r);
        const REAL am = (1.0 - w) * (v3 - v1) / (2.0 * dx) + w * (3.0 * v2 - 4.0 * v1 + v0) / (2.0 * dx);
        return v2 - dx * am;
    }
}

/* per-axis WENO estimate with the reference's branch order q==0, q==1, q==nc,
 * q==nc-1, interior: Grid3Drn.h:3085-3153 (k), :3200-3330 (j), :3333-3460 (i) */
static inline REAL FN(axisw_)(const REAL *tt, size_t n, size_t q, size_t nc, size_t st, REAL dx) {
    REAL a, t;
    if (q == 0) {
        a = tt[n + st];
    } else if (q == 1) {
        a = FN(weno3_)(0.0, tt[n - st], tt[n], tt[n + st], tt[n + 2 * st], dx, 1);
        t = tt[n - st];
        a = a < t ? a : t;
    } else if (q == nc) {
        a = tt[n - st];
    } else if (q == nc - 1) {
        a = FN(weno3_)(tt[n - 2 * st], tt[n - st], tt[n], tt[n + st], 0.0, dx, 0);
        t = tt[n + st];
        a = a < t ? a : t;
    } else {
        a = FN(weno3_)(tt[n - 2 * st], tt[n - st], tt[n], tt[n + st], tt[n + 2 * st], dx, 1);
        t = FN(weno3_)(tt[n - 2 * st], tt[n - st], tt[n], tt[n + st], tt[n + 2 * st], dx, 0);
        a = a < t ? a : t;
    }
    return a;
}

static inline void FN(update_)(FN(grid_) * g, size_t i, size_t j, size_t k, int weno) {
    const size_t n = NIDX(g, i, j, k);
    const size_t sj = g->nx1, sk = g->nx1 * g->ny1;
    REAL a1, a2, a3;
    if (!weno) { /* Grid3Drn.h:2902-2959 */
        a1 = FN(axis1_)(g->tt, n, k, g->nz1 - 1, sk);
        a2 = FN(axis1_)(g->tt, n, j, g->ny1 - 1, sj);
        a3 = FN(axis1_)(g->tt, n, i, g->nx1 - 1, 1);
    } else { /* Grid3Drn.h:3078-3484 */
        a1 = FN(axisw_)(g->tt, n, k, g->nz1 - 1, sk, g->dx);
        a2 = FN(axisw_)(g->tt, n, j, g->ny1 - 1, sj, g->dx);
        a3 = FN(axisw_)(g->tt, n, i, g->nx1 - 1, 1, g->dx);
    }
    FN(cascade_)(g, n, a1, a2, a3);
}

/* ---- sweep / sweep_weno3: Grid3Drn.h:2816-2899 / :2962-3044 ---------------
 * eight lexicographic passes, i fastest; directions in the order
 * +++, -++, +-+, --+, ++-, -+-, +--, ---  (sign of i, j, k).
 * order == 0: the reference's nested loops.
 * order == 1: the same eight passes visited by diagonal level (i'+j'+k' in the
 *             direction-oriented indices); any topological order of the
 *             Gauss-Seidel dependency DAG gives the same field (SURVEY section 6,
 *             Grid3Drn_OpenCL.h:839-848) -- checked bit-for-bit in tests. */
static void FN(sweep_)(FN(grid_) * g, const unsigned char *frozen, int weno, int order) {
    const long nx = (long)g->nx1, ny = (long)g->ny1, nz = (long)g->nz1;
    for (int d = 0; d < 8; ++d) {
        const int ri = d & 1, rj = (d >> 1) & 1, rk = (d >> 2) & 1;
        if (order == 0) {
            for (long kk = 0; kk < nz; ++kk) {
                const long k = rk ? nz - 1 - kk : kk;
                for (long jj = 0; jj < ny; ++jj) {
                    const long j = rj ? ny - 1 - jj : jj;
                    for (long ii = 0; ii < nx; ++ii) {
                        const long i = ri ? nx - 1 - ii : ii;
                        if (!frozen[NIDX(g, i, j, k)]) FN(update_)(g, (size_t)i, (size_t)j, (size_t)k, weno);
                    }
                }
            }
        } else {
            for (long lvl = 0; lvl <= nx + ny + nz - 3; ++lvl) {
                for (long kk = 0; kk < nz; ++kk) {
                    for (long jj = 0; jj < ny; ++jj) {
                        const long ii = lvl - kk - jj;
                        if (ii < 0 || ii >= nx) continue;
                        const long i = ri ? nx - 1 - ii : ii;
                        const long j = rj ? ny - 1 - jj : jj;
                        const long k = rk ? nz - 1 - kk : kk;
                        if (!frozen[NIDX(g, i, j, k)]) FN(update_)(g, (size_t)i, (size_t)j, (size_t)k, weno);
                    }
                }
            }
        }
    }
}

/* ---- initFSM: Grid3Drn.h:3487-3556 ---------------------------------------
 * node coordinates are min + idx*d evaluated in REAL (buildGridNodes,
 * Grid3Drn.h:391-399); node match is |d| < small = 1e-4 per axis (Node3Dn.h:147-149,
 * ttcr_t.h:42); the first matching node in index order wins (":3495-3496");
 * off-node sources use getCellNo (Grid3Drn.h:207-215) and skip the cell's
 * lower corner node (":3541"). */
static REAL FN(dist_)(REAL x, REAL y, REAL z, REAL px, REAL py, REAL pz) {
    /* Node3Dn.h:138-140 */
    return sqrt((x - px) * (x - px) + (y - py) * (y - py) + (z - pz) * (z - pz));
}

static void FN(init_fsm_)(FN(grid_) * g, REAL xmin, REAL ymin, REAL zmin, const REAL *tx, const REAL *t0,
                         size_t ntx, unsigned char *frozen, int npts) {
    const long ncx = (long)g->nx1 - 1, ncy = (long)g->ny1 - 1, ncz = (long)g->nz1 - 1;
    const REAL dx = g->dx;
    const double small = 1.e-4, small2 = small * small;
    const REAL xmax = xmin + ncx * dx, ymax = ymin + ncy * dx, zmax = zmin + ncz * dx; /* Grid3Drn.h:73 */
    for (size_t n = 0; n < ntx; ++n) {
        const REAL px = tx[3 * n], py = tx[3 * n + 1], pz = tx[3 * n + 2];
        long fi = -1, fj = -1, fk = -1;
        /* first node (k outer, j, i inner) within `small` on all three axes.  The
         * per-axis tests are independent, so the first match in index order is the
         * first matching index on each axis. */
        for (long k = 0; k <= ncz && fk < 0; ++k) { REAL z = zmin + k * dx; if (fabs(z - pz) < small) fk = k; }
        for (long j = 0; j <= ncy && fj < 0; ++j) { REAL y = ymin + j * dx; if (fabs(y - py) < small) fj = j; }
        for (long i = 0; i <= ncx && fi < 0; ++i) { REAL x = xmin + i * dx; if (fabs(x - px) < small) fi = i; }
        long i, j, k, lo;
        int on_node = (fi >= 0 && fj >= 0 && fk >= 0);
        if (on_node) {
            i = fi; j = fj; k = fk; lo = npts;
            const size_t nn = NIDX(g, i, j, k);
            g->tt[nn] = t0[n];
            frozen[nn] = 1;
        } else {
            REAL x = xmax - px < small2 ? xmax - .5 * dx : px;
            REAL y = ymax - py < small2 ? ymax - .5 * dx : py;
            REAL z = zmax - pz < small2 ? zmax - .5 * dx : pz;
            i = (long)(unsigned)(small2 + (x - xmin) / dx);
            j = (long)(unsigned)(small2 + (y - ymin) / dx);
            k = (long)(unsigned)(small2 + (z - zmin) / dx);
            lo = npts - 1;
        }
        for (long kk = k - lo; kk <= k + npts; ++kk) {
            if (kk < 0 || kk > ncz) continue;
            for (long jj = j - lo; jj <= j + npts; ++jj) {
                if (jj < 0 || jj > ncy) continue;
                for (long ii = i - lo; ii <= i + npts; ++ii) {
                    if (ii < 0 || ii > ncx || (ii == i && jj == j && kk == k)) continue;
                    const size_t nnn = NIDX(g, ii, jj, kk);
                    REAL X = xmin + ii * dx, Y = ymin + jj * dx, Z = zmin + kk * dx;
                    REAL t = t0[n] + FN(dist_)(X, Y, Z, px, py, pz) * g->s[nnn];
                    g->tt[nnn] = t;
                    frozen[nnn] = 1;
                }
            }
        }
    }
}

/* ---- driver: Grid3Drnfs::raytrace, Grid3Drnfs.h:84-155 (== Grid3Drcfs.h:175-247)
 * returns 0, or 1 if a Tx point is outside the grid (checkPts, Grid3Drn.h:771-790).
 * eps is the per-node tolerance; the constructor turns it into an L1 threshold
 * eps * N in REAL (Grid3Drnfs.h:49). */
int FN(fsmo_solve)(size_t ncx, size_t ncy, size_t ncz, REAL dx, REAL xmin, REAL ymin, REAL zmin, REAL eps,
                   int maxit, int weno, const REAL *s_node, const REAL *tx, const REAL *t0, size_t ntx,
                   REAL *tt, int *niter_out, int *niterw_out, int order) {
    FN(grid_) g = {ncx + 1, ncy + 1, ncz + 1, dx, tt, s_node};
    const size_t N = g.nx1 * g.ny1 * g.nz1;
    const REAL xmax = xmin + ncx * dx, ymax = ymin + ncy * dx, zmax = zmin + ncz * dx;
    for (size_t n = 0; n < ntx; ++n) {
        if (tx[3 * n] < xmin || tx[3 * n] > xmax || tx[3 * n + 1] < ymin || tx[3 * n + 1] > ymax ||
            tx[3 * n + 2] < zmin || tx[3 * n + 2] > zmax)
            return 1;
    }
    REAL epsilon = eps;
    epsilon *= (REAL)N;
    for (size_t n = 0; n < N; ++n) tt[n] = REAL_MAX; /* Node3Dn.h:103-105 */
    unsigned char *frozen = (unsigned char *)calloc(N, 1);
    REAL *times = (REAL *)malloc(N * sizeof(REAL));
    FN(init_fsm_)(&g, xmin, ymin, zmin, tx, t0, ntx, frozen, weno ? 2 : 1);
    for (size_t n = 0; n < N; ++n) times[n] = tt[n];
    int niter = 0, niterw = 0;
    REAL change = REAL_MAX;
    while (change >= epsilon && niter < maxit) {
        FN(sweep_)(&g, frozen, 0, order);
        change = 0.0;
        for (size_t n = 0; n < N; ++n) {
            REAL dt = times[n] - tt[n];
            dt = dt < 0 ? -dt : dt;
            change += dt;
            times[n] = tt[n];
        }
        niter++;
    }
    if (weno) {
        change = REAL_MAX;
        while (change >= epsilon && niterw < maxit) {
            FN(sweep_)(&g, frozen, 1, order);
            change = 0.0;
            for (size_t n = 0; n < N; ++n) {
                REAL dt = times[n] - tt[n];
                dt = dt < 0 ? -dt : dt;
                change += dt;
                times[n] = tt[n];
            }
            niterw++;
        }
    }
    *niter_out = niter;
    *niterw_out = niterw;
    free(frozen);
    free(times);
    return 0;
}

/* ---- receiver traveltimes: Grid3Drn::getTraveltime, Grid3Drn.h:794-930 -----
 * trilinear interpolation with the on-node / edge / face special cases at
 * tolerance small2 = 1e-8; interpolation order z, then y, then x. */
void FN(fsmo_interp)(size_t ncx, size_t ncy, size_t ncz, REAL dx, REAL xmin, REAL ymin, REAL zmin,
                     const REAL *tt, const REAL *rx, size_t nrx, REAL *out) {
    const size_t nnx = ncx + 1, nny = ncy + 1;
    (void)ncz;
    const double small2 = 1.e-8;
    for (size_t r = 0; r < nrx; ++r) {
        const REAL px = rx[3 * r], py = rx[3 * r + 1], pz = rx[3 * r + 2];
        const size_t i = (unsigned)(small2 + (px - xmin) / dx);
        const size_t j = (unsigned)(small2 + (py - ymin) / dx);
        const size_t k = (unsigned)(small2 + (pz - zmin) / dx);
        const int onx = fabs(px - (xmin + i * dx)) < small2;
        const int ony = fabs(py - (ymin + j * dx)) < small2;
        const int onz = fabs(pz - (zmin + k * dx)) < small2;
#define TT(ii, jj, kk) tt[(((kk)) * nny + (jj)) * nnx + (ii)]
        REAL v;
        if (onx && ony && onz) {
            v = TT(i, j, k);
        } else if (onx && ony) {
            REAL t1 = TT(i, j, k), t2 = TT(i, j, k + 1);
            REAL w1 = (zmin + (k + 1) * dx - pz) / dx, w2 = (pz - (zmin + k * dx)) / dx;
            v = t1 * w1 + t2 * w2;
        } else if (onx && onz) {
            REAL t1 = TT(i, j, k), t2 = TT(i, j + 1, k);
            REAL w1 = (ymin + (j + 1) * dx - py) / dx, w2 = (py - (ymin + j * dx)) / dx;
            v = t1 * w1 + t2 * w2;
        } else if (ony && onz) {
            REAL t1 = TT(i, j, k), t2 = TT(i + 1, j, k);
            REAL w1 = (xmin + (i + 1) * dx - px) / dx, w2 = (px - (xmin + i * dx)) / dx;
            v = t1 * w1 + t2 * w2;
        } else if (onx) {
            REAL t1 = TT(i, j, k), t2 = TT(i, j, k + 1), t3 = TT(i, j + 1, k), t4 = TT(i, j + 1, k + 1);
            REAL w1 = (zmin + (k + 1) * dx - pz) / dx, w2 = (pz - (zmin + k * dx)) / dx;
            t1 = t1 * w1 + t2 * w2;
            t2 = t3 * w1 + t4 * w2;
            w1 = (ymin + (j + 1) * dx - py) / dx;
            w2 = (py - (ymin + j * dx)) / dx;
            v = t1 * w1 + t2 * w2;
        } else if (ony) {
            REAL t1 = TT(i, j, k), t2 = TT(i, j, k + 1), t3 = TT(i + 1, j, k), t4 = TT(i + 1, j, k + 1);
            REAL w1 = (zmin + (k + 1) * dx - pz) / dx, w2 = (pz - (zmin + k * dx)) / dx;
            t1 = t1 * w1 + t2 * w2;
            t2 = t3 * w1 + t4 * w2;
            w1 = (xmin + (i + 1) * dx - px) / dx;
            w2 = (px - (xmin + i * dx)) / dx;
            v = t1 * w1 + t2 * w2;
        } else if (onz) {
            REAL t1 = TT(i, j, k), t2 = TT(i, j + 1, k), t3 = TT(i + 1, j, k), t4 = TT(i + 1, j + 1, k);
            REAL w1 = (ymin + (j + 1) * dx - py) / dx, w2 = (py - (ymin + j * dx)) / dx;
            t1 = t1 * w1 + t2 * w2;
            t2 = t3 * w1 + t4 * w2;
            w1 = (xmin + (i + 1) * dx - px) / dx;
            w2 = (px - (xmin + i * dx)) / dx;
            v = t1 * w1 + t2 * w2;
        } else {
            REAL t1 = TT(i, j, k), t2 = TT(i, j, k + 1), t3 = TT(i, j + 1, k), t4 = TT(i, j + 1, k + 1);
            REAL t5 = TT(i + 1, j, k), t6 = TT(i + 1, j, k + 1), t7 = TT(i + 1, j + 1, k), t8 = TT(i + 1, j + 1, k + 1);
            REAL w1 = (zmin + (k + 1) * dx - pz) / dx, w2 = (pz - (zmin + k * dx)) / dx;
            t1 = t1 * w1 + t2 * w2;
            t2 = t3 * w1 + t4 * w2;
            t3 = t5 * w1 + t6 * w2;
            t4 = t7 * w1 + t8 * w2;
            w1 = (ymin + (j + 1) * dx - py) / dx;
            w2 = (py - (ymin + j * dx)) / dx;
            t1 = t1 * w1 + t2 * w2;
            t2 = t3 * w1 + t4 * w2;
            w1 = (xmin + (i + 1) * dx - px) / dx;
            w2 = (px - (xmin + i * dx)) / dx;
            v = t1 * w1 + t2 * w2;
        }
#undef TT
        out[r] = v;
    }
}

/* ---- receiver traveltimes along raypaths: Grid3Drn::getTraveltimeFromRaypath, Grid3Drn.h:1103-1243 -----------
 * with its helpers grad (4th-order centred differences of interpolated traveltimes, :1032-1100; note the x axis
 * starts its stencil at pt.x - dx, the other two at pt - d/2), getIJK (:239-243) and computeSlowness (:2451-2676,
 * processVel == false: slowness interpolated with Interpolator::linear / bilinear / trilinear, Interpolator.h:37-85).
 * Mixed precision is the reference's: REAL variables, double literals. */
static REAL FN(tt_at_)(size_t ncx, size_t ncy, size_t ncz, REAL dx, REAL xmin, REAL ymin, REAL zmin, const REAL *tt,
                       REAL px, REAL py, REAL pz) {
    REAL rx[3] = {px, py, pz}, out;
    FN(fsmo_interp)(ncx, ncy, ncz, dx, xmin, ymin, zmin, tt, rx, 1, &out);
    return out;
}

static REAL FN(slow_at_)(size_t ncx, size_t ncy, size_t ncz, REAL dx, REAL xmin, REAL ymin, REAL zmin, const REAL *sl,
                         REAL px, REAL py, REAL pz, int pv) {
    const size_t nnx = ncx + 1, nny = ncy + 1, nnz = ncz + 1;
    const double small = 1.e-4, small2 = small * small;
    ptrdiff_t onX = -1, onY = -1, onZ = -1;
    /* the reference scans all nodes of an axis for |p - (min + n d)| < small2; only the nodes around p can match */
    {
        const ptrdiff_t c = (ptrdiff_t)floor((double)(px - xmin) / (double)dx);
        for (ptrdiff_t n = c - 1 < 0 ? 0 : c - 1; n <= c + 2 && n < (ptrdiff_t)nnx; ++n)
            if (FABS(px - (xmin + n * dx)) < small2) { onX = n; break; }
    }
    {
        const ptrdiff_t c = (ptrdiff_t)floor((double)(py - ymin) / (double)dx);
        for (ptrdiff_t n = c - 1 < 0 ? 0 : c - 1; n <= c + 2 && n < (ptrdiff_t)nny; ++n)
            if (FABS(py - (ymin + n * dx)) < small2) { onY = n; break; }
    }
    {
        const ptrdiff_t c = (ptrdiff_t)floor((double)(pz - zmin) / (double)dx);
        for (ptrdiff_t n = c - 1 < 0 ? 0 : c - 1; n <= c + 2 && n < (ptrdiff_t)nnz; ++n)
            if (FABS(pz - (zmin + n * dx)) < small2) { onZ = n; break; }
    }
#define SN0(ii, jj, kk) sl[((size_t)(kk) * nny + (size_t)(jj)) * nnx + (size_t)(ii)]
/* processVel (interp_vel): the VELOCITIES of the nodes are interpolated and the result is inverted, Grid3Drn.h:2489-2669 */
#define SN(ii, jj, kk) (pv ? (REAL)(1.0 / SN0(ii, jj, kk)) : SN0(ii, jj, kk))
#define RET(e) do { const REAL r_ = (e); return pv ? (REAL)(1.0 / r_) : r_; } while (0)
    if (onX != -1 && onY != -1 && onZ != -1) return SN0(onX, onY, onZ);
    const unsigned i = (unsigned)(small + (px - xmin) / dx);
    const unsigned j = (unsigned)(small + (py - ymin) / dx);
    const unsigned k = (unsigned)(small + (pz - zmin) / dx);
    REAL x[3], y[3], z[3], s[8];
    if (onX != -1 && onY != -1) {
        s[0] = SN(onX, onY, k); s[1] = SN(onX, onY, k + 1);
        x[0] = pz; x[1] = zmin + k * dx; x[2] = zmin + (k + 1) * dx;
        RET((s[0] * (x[2] - x[0]) + s[1] * (x[0] - x[1])) / (x[2] - x[1]));
    } else if (onX != -1 && onZ != -1) {
        s[0] = SN(onX, j, onZ); s[1] = SN(onX, j + 1, onZ);
        x[0] = py; x[1] = ymin + j * dx; x[2] = ymin + (j + 1) * dx;
        RET((s[0] * (x[2] - x[0]) + s[1] * (x[0] - x[1])) / (x[2] - x[1]));
    } else if (onY != -1 && onZ != -1) {
        s[0] = SN(i, onY, onZ); s[1] = SN(i + 1, onY, onZ);
        x[0] = px; x[1] = xmin + i * dx; x[2] = xmin + (i + 1) * dx;
        RET((s[0] * (x[2] - x[0]) + s[1] * (x[0] - x[1])) / (x[2] - x[1]));
    } else if (onX != -1 || onY != -1 || onZ != -1) {
        if (onX != -1) {
            s[0] = SN(onX, j, k); s[1] = SN(onX, j, k + 1); s[2] = SN(onX, j + 1, k); s[3] = SN(onX, j + 1, k + 1);
            x[0] = py; y[0] = pz; x[1] = ymin + j * dx; y[1] = zmin + k * dx; x[2] = ymin + (j + 1) * dx; y[2] = zmin + (k + 1) * dx;
        } else if (onY != -1) {
            s[0] = SN(i, onY, k); s[1] = SN(i, onY, k + 1); s[2] = SN(i + 1, onY, k); s[3] = SN(i + 1, onY, k + 1);
            x[0] = px; y[0] = pz; x[1] = xmin + i * dx; y[1] = zmin + k * dx; x[2] = xmin + (i + 1) * dx; y[2] = zmin + (k + 1) * dx;
        } else {
            s[0] = SN(i, j, onZ); s[1] = SN(i, j + 1, onZ); s[2] = SN(i + 1, j, onZ); s[3] = SN(i + 1, j + 1, onZ);
            x[0] = px; y[0] = py; x[1] = xmin + i * dx; y[1] = ymin + j * dx; x[2] = xmin + (i + 1) * dx; y[2] = ymin + (j + 1) * dx;
        }
        RET((s[0] * (x[2] - x[0]) * (y[2] - y[0]) + s[1] * (x[2] - x[0]) * (y[0] - y[1]) + s[2] * (x[0] - x[1]) * (y[2] - y[0]) +
             s[3] * (x[0] - x[1]) * (y[0] - y[1])) /
            ((x[2] - x[1]) * (y[2] - y[1])));
    }
    s[0] = SN(i, j, k); s[1] = SN(i, j, k + 1); s[2] = SN(i, j + 1, k); s[3] = SN(i, j + 1, k + 1);
    s[4] = SN(i + 1, j, k); s[5] = SN(i + 1, j, k + 1); s[6] = SN(i + 1, j + 1, k); s[7] = SN(i + 1, j + 1, k + 1);
    x[0] = px; y[0] = py; z[0] = pz;
    x[1] = xmin + i * dx; y[1] = ymin + j * dx; z[1] = zmin + k * dx;
    x[2] = xmin + (i + 1) * dx; y[2] = ymin + (j + 1) * dx; z[2] = zmin + (k + 1) * dx;
#undef SN
#undef SN0
    RET((s[0] * (x[2] - x[0]) * (y[2] - y[0]) * (z[2] - z[0]) + s[1] * (x[2] - x[0]) * (y[2] - y[0]) * (z[0] - z[1]) +
         s[2] * (x[2] - x[0]) * (y[0] - y[1]) * (z[2] - z[0]) + s[3] * (x[2] - x[0]) * (y[0] - y[1]) * (z[0] - z[1]) +
         s[4] * (x[0] - x[1]) * (y[2] - y[0]) * (z[2] - z[0]) + s[5] * (x[0] - x[1]) * (y[2] - y[0]) * (z[0] - z[1]) +
         s[6] * (x[0] - x[1]) * (y[0] - y[1]) * (z[2] - z[0]) + s[7] * (x[0] - x[1]) * (y[0] - y[1]) * (z[0] - z[1])) /
        ((x[2] - x[1]) * (y[2] - y[1]) * (z[2] - z[1])));
#undef RET
}

/* Grid3Drn::computeSlowness at n points (x, y, z rows); used by the tests of Grid3d.get_s0 */
void FN(fsmo_slowness_at)(size_t ncx, size_t ncy, size_t ncz, REAL dx, REAL xmin, REAL ymin, REAL zmin, const REAL *sl,
                          const REAL *pts, size_t n, int interp_vel, REAL *out) {
    for (size_t r = 0; r < n; ++r)
        out[r] = FN(slow_at_)(ncx, ncy, ncz, dx, xmin, ymin, zmin, sl, pts[3 * r], pts[3 * r + 1], pts[3 * r + 2], interp_vel);
}

/* one axis of grad(): stencil points p1..p4 (first = p - off), shifted inwards at the grid faces */
static void FN(stencil_)(REAL p, REAL off, REAL d, REAL lo, REAL hi, REAL q[4]) {
    REAL p1 = p - off;
    REAL p2 = p1 + 0.5 * d, p3 = p1 + 1.5 * d, p4 = p1 + 2.0 * d;
    if (p1 <= lo) {
        p1 = lo; p2 = p1 + 0.5 * d; p3 = p1 + 1.5 * d; p4 = p1 + 2.0 * d;
    } else if (p4 >= hi) {
        p4 = hi; p3 = p4 - 0.5 * d; p2 = p4 - 1.5 * d; p1 = p4 - 2.0 * d;
    }
    q[0] = p1; q[1] = p2; q[2] = p3; q[3] = p4;
}

static int FN(sgn_)(REAL v) { return v > 0 ? 1 : (v < 0 ? -1 : 0); }   /* boost::math::sign */

/* returns 0, or 1 when a ray leaves the grid (the reference throws), 2 when it does not reach a source */
/* rays == NULL: traveltimes only.  Otherwise Grid3Drn::getRaypath (Grid3Drn.h:1339-1500, the same walk with
 * r_data.push_back): ray r gets npts[r] points, the first `cap` of which are stored at rays + 3 * cap * r. */
int FN(fsmo_raypaths)(size_t ncx, size_t ncy, size_t ncz, REAL dx, REAL xmin, REAL ymin, REAL zmin, const REAL *tt,
                      const REAL *sl, const REAL *tx, const REAL *t0, size_t ntx, const REAL *rx, size_t nrx, REAL *out,
                      REAL *rays, size_t *npts, size_t cap, int interp_vel) {
    const double small2 = 1.e-8;
    const REAL xmax = xmin + ncx * dx, ymax = ymin + ncy * dx, zmax = zmin + ncz * dx;   /* Grid3Drn ctor, :63-65 */
    const REAL k1 = 1. / 24., k2 = 9. / 8.;
    const REAL maxDist = SQRT(dx * dx + dx * dx + dx * dx);
#define TTAT(a, b, c) FN(tt_at_)(ncx, ncy, ncz, dx, xmin, ymin, zmin, tt, a, b, c)
#define SLAT(a, b, c) FN(slow_at_)(ncx, ncy, ncz, dx, xmin, ymin, zmin, sl, a, b, c, interp_vel)
#define DIST(ax, ay, az, bx, by, bz) SQRT(((ax) - (bx)) * ((ax) - (bx)) + ((ay) - (by)) * ((ay) - (by)) + ((az) - (bz)) * ((az) - (bz)))
    for (size_t r = 0; r < nrx; ++r) {
        const REAL Rx = rx[3 * r], Ry = rx[3 * r + 1], Rz = rx[3 * r + 2];
        REAL ttr = 0.0;
        int done = 0;
        size_t np_ = 0;
#define PUSH(a, b, c) do { if (rays && np_ < cap) { REAL *q_ = rays + 3 * (cap * r + np_); q_[0] = (a); q_[1] = (b); q_[2] = (c); } ++np_; } while (0)
        PUSH(Rx, Ry, Rz);
        for (size_t ns = 0; ns < ntx; ++ns)
            if (Rx == tx[3 * ns] && Ry == tx[3 * ns + 1] && Rz == tx[3 * ns + 2]) { ttr = t0[ns]; done = 1; break; }
        if (done) { out[r] = ttr; if (npts) npts[r] = np_; continue; }
        REAL px = Rx, py = Ry, pz = Rz;   /* prev_pt */
        REAL cx = Rx, cy = Ry, cz = Rz;   /* curr_pt */
        REAL s1 = SLAT(cx, cy, cz), s2;
        int reached = 0;
        size_t guard = 0;
        const size_t guard_max = 16 * (ncx + ncy + ncz) + 1024;
        while (!reached) {
            if (++guard > guard_max) return 2;
            REAL q[4], gx, gy, gz;
            FN(stencil_)(cx, dx, dx, xmin, xmax, q);          /* x: first point at pt.x - dx (sic) */
            gx = (k1 * TTAT(q[0], cy, cz) - k2 * TTAT(q[1], cy, cz) + k2 * TTAT(q[2], cy, cz) - k1 * TTAT(q[3], cy, cz)) / dx;
            FN(stencil_)(cy, dx / 2.0, dx, ymin, ymax, q);
            gy = (k1 * TTAT(cx, q[0], cz) - k2 * TTAT(cx, q[1], cz) + k2 * TTAT(cx, q[2], cz) - k1 * TTAT(cx, q[3], cz)) / dx;
            FN(stencil_)(cz, dx / 2.0, dx, zmin, zmax, q);
            gz = (k1 * TTAT(cx, cy, q[0]) - k2 * TTAT(cx, cy, q[1]) + k2 * TTAT(cx, cy, q[2]) - k1 * TTAT(cx, cy, q[3])) / dx;
            gx *= (REAL)-1.0; gy *= (REAL)-1.0; gz *= (REAL)-1.0;
            {
                /* one step against the gradient, to the next grid plane */
                ptrdiff_t i = (ptrdiff_t)(small2 + (cx - xmin) / dx);
                ptrdiff_t j = (ptrdiff_t)(small2 + (cy - ymin) / dx);
                ptrdiff_t k = (ptrdiff_t)(small2 + (cz - zmin) / dx);
                REAL xp = xmin + dx * (i + (FN(sgn_)(gx) > 0.0 ? 1.0 : 0.0));
                REAL yp = ymin + dx * (j + (FN(sgn_)(gy) > 0.0 ? 1.0 : 0.0));
                REAL zp = zmin + dx * (k + (FN(sgn_)(gz) > 0.0 ? 1.0 : 0.0));
                if (FABS(xp - cx) < small2) xp += dx * FN(sgn_)(gx);
                if (FABS(yp - cy) < small2) yp += dx * FN(sgn_)(gy);
                if (FABS(zp - cz) < small2) zp += dx * FN(sgn_)(gz);
                const REAL ax = gx != 0.0 ? (xp - cx) / gx : REAL_MAX;
                const REAL ay = gy != 0.0 ? (yp - cy) / gy : REAL_MAX;
                const REAL az = gz != 0.0 ? (zp - cz) / gz : REAL_MAX;
                if (ax < ay && ax < az) {
                    cx += ax * gx; cy += ax * gy; cz += ax * gz; cx = xp;
                } else if (ay < az) {
                    cx += ay * gx; cy += ay * gy; cz += ay * gz; cy = yp;
                } else {
                    cx += az * gx; cy += az * gy; cz += az * gz; cz = zp;
                }
                if (cx < xmin || cx > xmax || cy < ymin || cy > ymax || cz < zmin || cz > zmax) return 1;
                s2 = SLAT(cx, cy, cz);
                ttr += 0.5 * (s1 + s2) * DIST(px, py, pz, cx, cy, cz);
                s1 = s2;
                px = cx; py = cy; pz = cz;
                PUSH(cx, cy, cz);
                /* are we close enough to one of the Tx points?  (the reference does not leave this loop early) */
                for (size_t ns = 0; ns < ntx; ++ns) {
                    const REAL Tx = tx[3 * ns], Ty = tx[3 * ns + 1], Tz = tx[3 * ns + 2];
                    const REAL dist = DIST(cx, cy, cz, Tx, Ty, Tz);
                    if (dist < maxDist) {
                        gx = Tx - cx; gy = Ty - cy; gz = Tz - cz;
                        /* the construction of pass 0 once more, towards Tx */
                        ptrdiff_t i2 = (ptrdiff_t)(small2 + (cx - xmin) / dx);
                        ptrdiff_t j2 = (ptrdiff_t)(small2 + (cy - ymin) / dx);
                        ptrdiff_t k2i = (ptrdiff_t)(small2 + (cz - zmin) / dx);
                        REAL xq = xmin + dx * (i2 + (FN(sgn_)(gx) > 0.0 ? 1.0 : 0.0));
                        REAL yq = ymin + dx * (j2 + (FN(sgn_)(gy) > 0.0 ? 1.0 : 0.0));
                        REAL zq = zmin + dx * (k2i + (FN(sgn_)(gz) > 0.0 ? 1.0 : 0.0));
                        if (FABS(xq - cx) < small2) xq += dx * FN(sgn_)(gx);
                        if (FABS(yq - cy) < small2) yq += dx * FN(sgn_)(gy);
                        if (FABS(zq - cz) < small2) zq += dx * FN(sgn_)(gz);
                        const REAL bx = gx != 0.0 ? (xq - cx) / gx : REAL_MAX;
                        const REAL by = gy != 0.0 ? (yq - cy) / gy : REAL_MAX;
                        const REAL bz = gz != 0.0 ? (zq - cz) / gz : REAL_MAX;
                        if (bx < by && bx < bz) {
                            cx += bx * gx; cy += bx * gy; cz += bx * gz; cx = xq;
                        } else if (by < bz) {
                            cx += by * gx; cy += by * gy; cz += by * gz; cy = yq;
                        } else {
                            cx += bz * gx; cy += bz * gy; cz += bz * gz; cz = zq;
                        }
                        if (DIST(cx, cy, cz, px, py, pz) > dist || (cx == Tx && cy == Ty && cz == Tz)) {
                            s2 = SLAT(Tx, Ty, Tz);
                            ttr += t0[ns] + 0.5 * (s1 + s2) * DIST(px, py, pz, Tx, Ty, Tz);
                            PUSH(Tx, Ty, Tz);
                        } else {
                            s2 = SLAT(cx, cy, cz);
                            ttr += 0.5 * (s1 + s2) * DIST(px, py, pz, cx, cy, cz);
                            PUSH(cx, cy, cz);
                            s1 = s2;
                            s2 = SLAT(Tx, Ty, Tz);
                            ttr += t0[ns] + 0.5 * (s1 + s2) * DIST(cx, cy, cz, Tx, Ty, Tz);
                            PUSH(Tx, Ty, Tz);
                        }
                        reached = 1;
                    }
                }
            }
        }
        out[r] = ttr;
        if (npts) npts[r] = np_;
#undef PUSH
    }
#undef TTAT
#undef SLAT
#undef DIST
    return 0;
}

int FN(fsmo_tt_from_rp)(size_t ncx, size_t ncy, size_t ncz, REAL dx, REAL xmin, REAL ymin, REAL zmin, const REAL *tt,
                        const REAL *sl, const REAL *tx, const REAL *t0, size_t ntx, const REAL *rx, size_t nrx, REAL *out) {
    return FN(fsmo_raypaths)(ncx, ncy, ncz, dx, xmin, ymin, zmin, tt, sl, tx, t0, ntx, rx, nrx, out, NULL, NULL, 0, 0);
}

#undef NIDX
#undef FN
#undef CAT
#undef CAT_
