/* TEST / BASELINE INFRASTRUCTURE ONLY -- never linked or loaded by the product (parcels_b200/).
 *
 * C + OpenMP restatement of ONE configuration of the hot path, used as the "strong CPU baseline" of
 * bench.py and pinned bit-for-bit against the NumPy oracle by tests/test_oracle_c.py:
 *   Kernel.execute([AdvectionRK4_3D, delete-on-error]) on a rectilinear A-grid with float64 node
 *   coordinates, float32 U,V,W of shape (T,Z,Y,X), a time axis, float32 particle positions.
 * Arithmetic follows the reference operation by operation (compile with -ffp-contract=off):
 *   _core/kernel.py:174-247, kernels/_advection.py:58-75, _core/field.py:250-405,
 *   _core/index_search.py:20-91, interpolators/_xinterpolators.py:78-190, _core/basegrid.py:259-278.
 */
#include <math.h>
#include <stdint.h>

typedef struct {
    const double *lon, *lat, *depth, *time; /* time: seconds since the interval start */
    int nx, ny, nz, nt;
    int spherical;
    double deg2m;
    const float *U, *V, *W;
} grid_t;

/* _search_1d_array: idx = clip(searchsorted(arr, x, 'left') - 1, 0, n-2), sentinels -2 / -1 */
static inline int search1d(const double* a, int n, double x, double* b) {
    int lo = 0, hi = n;
    while (lo < hi) { int m = (lo + hi) >> 1; if (a[m] < x) lo = m + 1; else hi = m; }
    if (x != x) lo = n;
    int i = lo - 1; if (i < 0) i = 0; if (i > n - 2) i = n - 2;
    *b = (x - a[i]) / (a[i + 1] - a[i]);
    if (x < a[0]) return -2;
    if (x > a[n - 1]) return -1;
    return i;
}

static inline long wrapi(int i, int n) { int w = i < 0 ? i + n : i; return w < 0 ? 0 : (w > n - 1 ? n - 1 : w); }
static inline long upi(int i, int n) { int w = i + 1; return w < 0 ? 0 : (w > n - 1 ? n - 1 : w); }

static inline double xlinear(const grid_t* g, const float* D, int ti, double tau, int zi, double zeta, int yi, double eta, int xi,
                             double xsi) {
    const long sY = g->nx, sZ = (long)g->nx * g->ny, sT = sZ * g->nz;
    const long ot[2] = {wrapi(ti, g->nt) * sT, upi(ti, g->nt) * sT}, oz[2] = {wrapi(zi, g->nz) * sZ, upi(zi, g->nz) * sZ};
    const long oy[2] = {wrapi(yi, g->ny) * sY, upi(yi, g->ny) * sY}, ox[2] = {wrapi(xi, g->nx), upi(xi, g->nx)};
    double r[8];
    for (int k = 0; k < 8; ++k) {
        const long o = oz[(k >> 2) & 1] + oy[(k >> 1) & 1] + ox[k & 1];
        const double d0 = (double)D[ot[0] + o];
        r[k] = tau > 0 ? d0 * (1 - tau) + (double)D[ot[1] + o] * tau : d0;
    }
    if (zeta > 0) for (int k = 0; k < 4; ++k) r[k] = r[k] * (1 - zeta) + r[4 + k] * zeta;
    return (1 - xsi) * (1 - eta) * r[0] + xsi * (1 - eta) * r[1] + (1 - xsi) * eta * r[2] + xsi * eta * r[3];
}

/* VectorField.eval with XLinear_Velocity; y_is_f32: the sampled position is the particle's own float32 array */
static inline void eval_uvw(const grid_t* g, double t, double z, double y, double x, int y_is_f32, int* state, int* ei, double* u,
                            double* v, double* w) {
    if (!(0 <= t && t <= g->time[g->nt - 1])) { *state = 70; *u = *v = *w = 0; return; }
    double tau, zeta, eta, xsi;
    const int ti = search1d(g->time, g->nt, t, &tau);
    const int zi = search1d(g->depth, g->nz, z, &zeta);
    const int yi = search1d(g->lat, g->ny, y, &eta);
    const int xi = search1d(g->lon, g->nx, x, &xsi);
    *ei = (int)((long)zi * ((long)(g->ny - 1) * (g->nx - 1)) + (long)yi * (g->nx - 1) + xi);
    int s = *state;
    if (xi == -1 || yi == -1 || zi == -1) s = s > 60 ? s : 60;
    if (zi == -2) s = s > 61 ? s : 61;
    double uu = xlinear(g, g->U, ti, tau, zi, zeta, yi, eta, xi, xsi);
    double vv = xlinear(g, g->V, ti, tau, zi, zeta, yi, eta, xi, xsi);
    if (g->spherical) {
        const double conv = y_is_f32 ? (double)((float)g->deg2m * cosf((float)y * (float)(3.14159265358979323846 / 180.0)))
                                     : g->deg2m * cos(y * (3.14159265358979323846 / 180.0));
        uu /= conv;
        vv /= g->deg2m;
    }
    double ww = xlinear(g, g->W, ti, tau, zi, zeta, yi, eta, xi, xsi);
    if (uu != uu || vv != vv || ww != ww) s = s > 51 ? s : 51;
    if (xi < 0 || yi < 0 || zi < 0) uu = vv = ww = 0;
    *state = s;
    *u = uu; *v = vv; *w = ww;
}

/* returns the number of particle-steps evaluated; particles with state 30 (Delete) are left for the caller to drop */
long advect_rk4_3d(const grid_t* g, long n, float* x, float* y, float* z, double* t, int* state, int* ei, double dt, double endtime) {
    long steps = 0;
    const int sign = dt > 0 ? 1 : -1;
#pragma omp parallel for schedule(static) reduction(+ : steps)
    for (long i = 0; i < n; ++i) {
        float px = x[i], py = y[i], pz = z[i];
        double pt = t[i];
        int st = 10, e = ei[i];
        for (;;) {
            const double tte = sign * (endtime - pt);
            if (!((st == 0 || st == 10) && tte >= 0)) break;
            const double h = sign == 1 ? fmax(fmin(dt, tte), 0.0) : fmin(fmax(dt, -tte), 0.0);
            ++steps;
            double u1, v1, w1, u2, v2, w2, u3, v3, w3, u4, v4, w4;
            eval_uvw(g, pt, pz, py, px, 1, &st, &e, &u1, &v1, &w1);
            eval_uvw(g, pt + 0.5 * h, (double)pz + w1 * 0.5 * h, (double)py + v1 * 0.5 * h, (double)px + u1 * 0.5 * h, 0, &st, &e, &u2, &v2, &w2);
            eval_uvw(g, pt + 0.5 * h, (double)pz + w2 * 0.5 * h, (double)py + v2 * 0.5 * h, (double)px + u2 * 0.5 * h, 0, &st, &e, &u3, &v3, &w3);
            eval_uvw(g, pt + h, (double)pz + w3 * h, (double)py + v3 * h, (double)px + u3 * h, 0, &st, &e, &u4, &v4, &w4);
            const float dx = (float)(0.0 + (u1 + 2 * u2 + 2 * u3 + u4) / 6 * h);
            const float dy = (float)(0.0 + (v1 + 2 * v2 + 2 * v3 + v4) / 6 * h);
            const float dz = (float)(0.0 + (w1 + 2 * w2 + 2 * w3 + w4) / 6 * h);
            if (st >= 50) st = 30; /* delete-on-error handler */
            if (st == 10 || st == 0) { px = px + dx; py = py + dy; pz = pz + dz; pt = pt + h; }
            if (st == 10 && pt == endtime) st = 1;
            if (st == 30) break;
        }
        x[i] = px; y[i] = py; z[i] = pz; t[i] = pt; state[i] = st; ei[i] = e;
    }
    return steps;
}
