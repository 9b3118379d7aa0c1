/* TEST INFRASTRUCTURE ONLY -- body of the plain-C restatement of the reference's 2-D rectilinear fast-sweeping path
 * (Grid2Drnfs), included twice by fsm2d_oracle.c (REAL = double, REAL = float).
 * It is the checker a CUDA implementation of that row will be tested against; nothing in the product uses it.
 * All file:line citations are relative to /root/reference/ttcr/.  Same arithmetic conventions as fsm_oracle_impl.h:
 * variables are REAL, literals are double, no contraction (-ffp-contract=off).
 *
 * Node index n = i * (ncz + 1) + j, z fastest (Grid2Drnfs.h buildGridNodes); node coordinates min + idx * d in REAL. */
#ifndef REAL
#error "define REAL, SFX, REAL_MAX, REAL_EPS before including"
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct {
    long ncx, ncz;
    REAL dx, dz;
    REAL *tt;
    const REAL *s;
} FN(g2_);
#define N2(g, i, j) ((size_t)(i) * (size_t)((g)->ncz + 1) + (size_t)(j))

/* first-order one-sided minimum along one axis (Grid2Drn.h:924-943) */
static inline REAL FN(ax1_)(const REAL *tt, size_t n, long q, long nc, size_t st) {
    if (q == 0) return tt[n + st];
    if (q == nc) return tt[n - st];
    REAL a = tt[n - st], t = tt[n + st];
    return a < t ? a : t;
}

/* third-order WENO one-sided estimate (Grid2Drn.h:1083-1091 forward, :1101-1109 backward) */
static inline REAL FN(w3_)(REAL v0, REAL v1, REAL v2, REAL v3, REAL v4, REAL d, int forward) {
    const REAL eps = REAL_EPS;
    REAL den = v3 - 2. * v2 + v1;
    den *= den;
    if (forward) {
        REAL num = v4 - 2. * v3 + v2;
        num *= num;
        const REAL r = (eps + num) / (eps + den);
        const REAL w = 1. / (1. + 2. * r * r);
        const REAL ap = (1. - w) * (v3 - v1) / (2. * d) + w * (-v4 + 4. * v3 - 3. * v2) / (2. * d);
        return v2 + d * ap;
    } else {
        REAL num = v2 - 2. * v1 + v0;
        num *= num;
        const REAL r = (eps + num) / (eps + den);
        const REAL w = 1. / (1. + 2. * r * r);
        const REAL am = (1. - w) * (v3 - v1) / (2. * d) + w * (3. * v2 - 4. * v1 + v0) / (2. * d);
        return v2 - d * am;
    }
}

/* per-axis WENO estimate, branch order q == 0, 1, nc, nc-1, interior (Grid2Drn.h:1080-1134 for i, :1136-1190 for j) */
static inline REAL FN(axw_)(const REAL *tt, size_t n, long q, long nc, size_t st, REAL d) {
    REAL a, t;
    if (q == 0) {
        a = tt[n + st];
    } else if (q == 1) {
        a = FN(w3_)(0.0, tt[n - st], tt[n], tt[n + st], tt[n + 2 * st], d, 1);
        t = tt[n - st];
        a = a < t ? a : t;
    } else if (q == nc) {
        a = tt[n - st];
    } else if (q == nc - 1) {
        a = FN(w3_)(tt[n - 2 * st], tt[n - st], tt[n], tt[n + st], 0.0, d, 0);
        t = tt[n + st];
        a = a < t ? a : t;
    } else {
        const REAL ap = FN(w3_)(tt[n - 2 * st], tt[n - st], tt[n], tt[n + st], tt[n + 2 * st], d, 1);
        const REAL am = FN(w3_)(tt[n - 2 * st], tt[n - st], tt[n], tt[n + st], tt[n + 2 * st], d, 0);
        a = am < ap ? am : ap;
    }
    return a;
}

/* the two local solvers: square cells (Grid2Drn.h:945-953) and dx != dz (:1041-1057) */
static inline void FN(solve_sq_)(FN(g2_) * g, size_t n, REAL a, REAL b, REAL fh) {
    REAL t;
    REAL d = a - b;
    d = d < 0 ? -d : d;
    if (d >= fh) t = (a < b ? a : b) + fh;
    else t = 0.5 * (a + b + sqrt(2. * fh * fh - (a - b) * (a - b)));
    if (t < g->tt[n]) g->tt[n] = t;
}
static inline void FN(solve_xz_)(FN(g2_) * g, size_t n, REAL a, REAL b) {
    const REAL dx = g->dx, dz = g->dz, s = g->s[n];
    REAL t;
    if (a < b && ((b - a) / dx) > s) {
        t = a + s * dx;
    } else if (a > b && ((a - b) / dz) > s) {
        t = b + s * dz;
    } else {
        REAL dx2 = dx * dx, dz2 = dz * dz, s2 = s * s;
        t = (b * dx2 + a * dz2) / (dx2 + dz2) +
            sqrt((2.0 * a * b * dx2 * dz2 - a * a * dx2 * dz2 - b * b * dx2 * dz2 + dx2 * dx2 * dz2 * s2 + dx2 * dz2 * dz2 * s2) /
                 ((dx2 + dz2) * (dx2 + dz2)));
    }
    if (t < g->tt[n]) g->tt[n] = t;
}

/* kind: 0 update_node (:920-955), 1 update_node45 (:957-1015), 2 update_node_xz (:1019-1058),
 *       3 update_node_weno3 (:1061-1214), 4 update_node_weno3_xz (:1217-1357) */
static inline void FN(update2_)(FN(g2_) * g, long i, long j, int kind) {
    const size_t st = (size_t)g->ncz + 1, n = N2(g, i, j);
    const long ncx = g->ncx, ncz = g->ncz;
    REAL a, b, t;
    switch (kind) {
        case 0:
            a = FN(ax1_)(g->tt, n, i, ncx, st);
            b = FN(ax1_)(g->tt, n, j, ncz, 1);
            FN(solve_sq_)(g, n, a, b, g->s[n] * g->dx);
            break;
        case 1: {   /* stencil rotated by pi/4: the diagonals (i+1,j+1)/(i-1,j-1) and (i+1,j-1)/(i-1,j+1); +MAX off the grid */
            const REAL M = REAL_MAX;
            REAL pp = (i != ncx && j != ncz) ? g->tt[n + st + 1] : M, mm = (i != 0 && j != 0) ? g->tt[n - st - 1] : M;
            REAL pm = (i != ncx && j != 0) ? g->tt[n + st - 1] : M, mp = (i != 0 && j != ncz) ? g->tt[n - st + 1] : M;
            if (i == 0) { a = pp; b = pm; }
            else if (i == ncx) { a = mm; b = mp; }
            else { a = pp; t = mm; a = a < t ? a : t; b = pm; t = mp; b = b < t ? b : t; }
            const REAL fh = 1.414213562373095 * g->s[n] * g->dx;
            FN(solve_sq_)(g, n, a, b, fh);
            break;
        }
        case 2:
            a = FN(ax1_)(g->tt, n, i, ncx, st);
            b = FN(ax1_)(g->tt, n, j, ncz, 1);
            FN(solve_xz_)(g, n, a, b);
            break;
        case 3:
            a = FN(axw_)(g->tt, n, i, ncx, st, g->dx);
            b = FN(axw_)(g->tt, n, j, ncz, 1, g->dx);     /* (sic: dx on both axes, the scheme requires dx == dz) */
            FN(solve_sq_)(g, n, a, b, g->s[n] * g->dx);
            break;
        default:
            a = FN(axw_)(g->tt, n, i, ncx, st, g->dx);
            b = FN(axw_)(g->tt, n, j, ncz, 1, g->dz);
            FN(solve_xz_)(g, n, a, b);
    }
}

/* the four Gauss-Seidel passes, identical for all five node updates (Grid2Drn.h:713-752, :756-794, :797-835, :838-876,
 * :879-917): (i up, j up), (i down, j up), (i down, j down), (i up, j down); i is the outer loop */
static void FN(sweep2_)(FN(g2_) * g, const unsigned char *frozen, int kind) {
    static const int di[4] = {1, -1, -1, 1}, dj[4] = {1, 1, -1, -1};
    for (int d = 0; d < 4; ++d)
        for (long ii = 0; ii <= g->ncx; ++ii) {
            const long i = di[d] > 0 ? ii : g->ncx - ii;
            for (long jj = 0; jj <= g->ncz; ++jj) {
                const long j = dj[d] > 0 ? jj : g->ncz - jj;
                if (!frozen[N2(g, i, j)]) FN(update2_)(g, i, j, kind);
            }
        }
}

/* initFSM, Grid2Drn.h:1360-1419: on a node (|d| < small on both axes, first node in index order) the (2 npts + 1)^2 box
 * around it gets t0 + dist * mean(slowness of the node, slowness of the source node); off-node the box anchored at the
 * cell (getCellNo, :173-179) gets t0 + dist * slowness(node), the cell's own corner included (unlike 3-D) */
static void FN(init2_)(FN(g2_) * g, REAL xmin, REAL zmin, const REAL *tx, const REAL *t0, size_t ntx, unsigned char *frozen, int npts) {
    const double small = 1.e-4;
    const long ncx = g->ncx, ncz = g->ncz;
    const REAL dx = g->dx, dz = g->dz;
    const REAL xmax = xmin + ncx * dx, zmax = zmin + ncz * dz;
    for (size_t n = 0; n < ntx; ++n) {
        const REAL px = tx[2 * n], pz = tx[2 * n + 1];
        long fi = -1, fj = -1;
        for (long i = 0; i <= ncx && fi < 0; ++i) { REAL x = xmin + i * dx; if (fabs(x - px) < small) fi = i; }
        for (long j = 0; j <= ncz && fj < 0; ++j) { REAL z = zmin + j * dz; if (fabs(z - pz) < small) fj = j; }
        if (fi >= 0 && fj >= 0) {
            const size_t nn = N2(g, fi, fj);
            g->tt[nn] = t0[n];
            frozen[nn] = 1;
            for (long ii = fi - npts; ii <= fi + npts; ++ii) {
                if (ii < 0 || ii > ncx) continue;
                for (long jj = fj - npts; jj <= fj + npts; ++jj) {
                    if (jj < 0 || jj > ncz || (ii == fi && jj == fj)) continue;
                    const size_t nnn = N2(g, ii, jj);
                    const REAL X = xmin + ii * dx, Z = zmin + jj * dz;
                    const REAL dist = sqrt((X - px) * (X - px) + (Z - pz) * (Z - pz));
                    g->tt[nnn] = t0[n] + dist * 0.5 * (g->s[nnn] + g->s[nn]);
                    frozen[nnn] = 1;
                }
            }
        } else {
            const REAL x = xmax - px < small ? xmax - .5 * dx : px;
            const REAL z = zmax - pz < small ? zmax - .5 * dz : pz;
            const long i = (long)(unsigned)(small + (x - xmin) / dx), j = (long)(unsigned)(small + (z - zmin) / dz);
            for (long ii = i - (npts - 1); ii <= i + npts; ++ii) {
                if (ii < 0 || ii > ncx) continue;
                for (long jj = j - (npts - 1); jj <= j + npts; ++jj) {
                    if (jj < 0 || jj > ncz) continue;
                    const size_t nnn = N2(g, ii, jj);
                    const REAL X = xmin + ii * dx, Z = zmin + jj * dz;
                    const REAL dist = sqrt((X - px) * (X - px) + (Z - pz) * (Z - pz));
                    g->tt[nnn] = t0[n] + dist * g->s[nnn];
                    frozen[nnn] = 1;
                }
            }
        }
    }
}

static REAL FN(l1_)(REAL *times, const REAL *tt, size_t N) {
    REAL change = 0.0;
    for (size_t n = 0; n < N; ++n) {
        REAL dt = times[n] - tt[n];
        dt = dt < 0 ? -dt : dt;
        change += dt;
        times[n] = tt[n];
    }
    return change;
}

/* Grid2Drnfs::raytrace, Grid2Drnfs.h:195-299.  eps is per node (the constructor multiplies it by the node count, :92).
 * tx: ntx x 2 (x, z).  Returns 0, or 1 if a Tx point is outside the grid (checkPts, Grid2Drn.h:333-342). */
int FN(fsmo2d_solve)(size_t ncx, size_t ncz, REAL dx, REAL dz, REAL xmin, REAL zmin, REAL eps, int maxit, int weno, int rotated,
                     const REAL *s_node, const REAL *tx, const REAL *t0, size_t ntx, REAL *tt, int *niter_out, int *niterw_out) {
    FN(g2_) g = {(long)ncx, (long)ncz, dx, dz, tt, s_node};
    const size_t N = (ncx + 1) * (ncz + 1);
    const REAL xmax = xmin + ncx * dx, zmax = zmin + ncz * dz;   /* Grid2Drn ctor */
    for (size_t n = 0; n < ntx; ++n)
        if (tx[2 * n] < xmin || tx[2 * n] > xmax || tx[2 * n + 1] < zmin || tx[2 * n + 1] > zmax) return 1;
    REAL epsilon = eps;
    epsilon *= (REAL)N;
    for (size_t n = 0; n < N; ++n) tt[n] = REAL_MAX;
    unsigned char *frozen = (unsigned char *)calloc(N, 1);
    REAL *times = (REAL *)malloc(N * sizeof(REAL));
    FN(init2_)(&g, xmin, zmin, tx, t0, ntx, frozen, weno ? 2 : 1);
    for (size_t n = 0; n < N; ++n) times[n] = tt[n];
    int niter = 0, niterw = 0;
    REAL change = REAL_MAX;
    const int square = dx == dz;
    if (weno) {
        while (change >= epsilon && niter < maxit) {
            FN(sweep2_)(&g, frozen, square ? 0 : 2);
            change = FN(l1_)(times, tt, N);
            niter++;
        }
        change = REAL_MAX;
        while (change >= epsilon && niterw < maxit) {
            FN(sweep2_)(&g, frozen, square ? 3 : 4);
            change = FN(l1_)(times, tt, N);
            niterw++;
        }
    } else {
        while (change >= epsilon && niter < maxit) {
            if (square) {
                FN(sweep2_)(&g, frozen, 0);
                if (rotated) FN(sweep2_)(&g, frozen, 1);
            } else {
                FN(sweep2_)(&g, frozen, 2);
            }
            change = FN(l1_)(times, tt, N);
            niter++;
        }
    }
    *niter_out = niter;
    *niterw_out = niterw;
    free(frozen);
    free(times);
    return 0;
}

/* Grid2Drcfs::setSlowness, Grid2Drcfs.h:99-138: node slowness = mean of the 1 / 2 / 4 adjacent cells (cell index
 * i * ncz + j); the order of the additions is the source's */
void FN(fsmo2d_cell_to_node)(const REAL *s, size_t nx, size_t nz, REAL *sn) {
    const size_t st = nz + 1;
    sn[0] = s[0];
    sn[nz] = s[nz - 1];
    sn[nx * st] = s[nz * (nx - 1)];
    sn[(nx + 1) * st - 1] = s[nx * nz - 1];
    for (size_t j = 1; j < nz; ++j) {
        sn[j] = 0.5 * (s[j] + s[j - 1]);
        sn[nx * st + j] = 0.5 * (s[nz * (nx - 1) + j] + s[nz * (nx - 1) + j - 1]);
    }
    for (size_t i = 1; i < nx; ++i) {
        sn[i * st] = 0.5 * (s[i * nz] + s[(i - 1) * nz]);
        sn[i * st + nz] = 0.5 * (s[(i + 1) * nz - 1] + s[i * nz - 1]);
    }
    for (size_t i = 1; i < nx; ++i)
        for (size_t j = 1; j < nz; ++j)
            sn[i * st + j] = 0.25 * (s[i * nz + j] + s[i * nz + j - 1] + s[(i - 1) * nz + j] + s[(i - 1) * nz + j - 1]);
}

/* Grid2Drn::getTraveltime, Grid2Drn.h:359-415: bilinear, x first then z, with the on-node and on-edge cases at
 * tolerance small = 1e-4 (getIJ, :186-189) */
void FN(fsmo2d_interp)(size_t ncx, size_t ncz, REAL dx, REAL dz, REAL xmin, REAL zmin, const REAL *tt, const REAL *rx, size_t nrx,
                       REAL *out) {
    const double small = 1.e-4;
    const size_t nnz = ncz + 1;
    (void)ncx;
    for (size_t r = 0; r < nrx; ++r) {
        const REAL px = rx[2 * r], pz = rx[2 * r + 1];
        const size_t i = (unsigned)(small + (px - xmin) / dx), j = (unsigned)(small + (pz - zmin) / dz);
        const int onx = fabs(px - (xmin + i * dx)) < small, onz = fabs(pz - (zmin + j * dz)) < small;
        REAL t;
        if (onx && onz) {
            t = tt[i * nnz + j];
        } else if (onx) {
            const REAL t1 = tt[i * nnz + j], t2 = tt[i * nnz + j + 1];
            const REAL w1 = (zmin + (j + 1) * dz - pz) / dz, w2 = (pz - (zmin + j * dz)) / dz;
            t = t1 * w1 + t2 * w2;
        } else if (onz) {
            const REAL t1 = tt[i * nnz + j], t2 = tt[(i + 1) * nnz + j];
            const REAL w1 = (xmin + (i + 1) * dx - px) / dx, w2 = (px - (xmin + i * dx)) / dx;
            t = t1 * w1 + t2 * w2;
        } else {
            REAL t1 = tt[i * nnz + j], t2 = tt[(i + 1) * nnz + j];
            const REAL t3 = tt[i * nnz + j + 1], t4 = tt[(i + 1) * nnz + j + 1];
            REAL w1 = (xmin + (i + 1) * dx - px) / dx, w2 = (px - (xmin + i * dx)) / dx;
            t1 = t1 * w1 + t2 * w2;
            t2 = t3 * w1 + t4 * w2;
            w1 = (zmin + (j + 1) * dz - pz) / dz;
            w2 = (pz - (zmin + j * dz)) / dz;
            t = t1 * w1 + t2 * w2;
        }
        out[r] = t;
    }
}

#undef N2
#undef FN
#undef CAT
#undef CAT_
