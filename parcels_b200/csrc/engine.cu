// engine.cu -- libparcels_b200.so: hand-written sm_100a CUDA for the Parcels hot path
//   ParticleSet.execute(AdvectionRK4 | AdvectionRK4_3D | AdvectionEE | AdvectionRK2[_3D]
//                       [+ DiffusionUniformKh] [+ delete-on-error])
// behind the C-ABI declared in include/parcels_b200.h.
//
// Design (DESIGN.md has the long version):
//  * one lane per particle; the WHOLE dt loop of Kernel.execute (reference _core/kernel.py:174-247)
//    runs in-kernel with the particle in registers, so particle state crosses HBM once per call;
//  * the 2x2x2x2 (t,z,y,x) corner block of U,V,W that brackets a particle (16 values / component,
//    reference interpolators/_xinterpolators.py:78-96) is kept in REGISTERS together with the cell
//    bounds; a particle stays in its cell for tens of steps, so the 4 RK stages of most steps need
//    no memory traffic at all -- the field is gathered from HBM only when a stage position leaves
//    the cached cell ("corner cache refill");
//  * all arithmetic reproduces NumPy's dtype promotion of the reference (f32 particle positions,
//    f32/f64 grid coordinates and data, f64 time) operation by operation, with FMA contraction OFF
//    (-fmad=false), so cell indices are bit-exact and trajectories agree to the last f32 ulp
//    (transcendentals aside);
//  * no tensor cores: this is a gather + fp64 lerp path, there is no contraction to map on tcgen05.
//
// This file is compiled ONLY for sm_100a.  There is no CPU fallback anywhere in the product.

#include <cuda_runtime.h>
#include <math_constants.h>

#include <climits>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/parcels_b200.h"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(PB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side descriptors
// ------------------------------------------------------------------------------------------------
struct GridDev {
    const void* lon;
    const void* lat;
    const void* depth;
    const double* time;  // seconds since interval start (time[0] == 0 after host normalisation)
    int nx, ny, nz, nt;  // node counts; nz == 0: grid has no Z axis
    int spherical;
    int pad_;
    double deg2m;
    double time_len;  // time[nt-1] - time[0]
    long long xdim, ydim, zdim;  // cell counts for ravel_index
};

struct FieldDev {
    const void* p[3];
    int T, Z, Y, X;            // data shape (shared by all components on an A-grid)
    long long sT, sZ, sY, sX;  // element strides; 0 for size-1 (never indexed) dims
};

struct ReportDev {
    unsigned long long particle_steps;
    unsigned long long n_error;
    unsigned long long n_deleted;
    long long first_error_iter;  // LLONG_MAX when none
    unsigned long long n_out_of_time;
    long long max_iters_done;
    unsigned long long cache_refills;
    int max_state;
    int pad_;
};

struct ParticlesDev {
    float *x, *y, *z, *dx, *dy, *dz;
    double* t;
    int* state;
    int* ei;
    long long* pid;
    long long n;
};

struct AdvectParams {
    GridDev g;
    FieldDev f;
    ParticlesDev P;
    int scheme, diffusion, delete_on_error, kh_spherical;
    double dt, endtime, kh_zonal, kh_meridional, kh_deg2m;
    unsigned long long seed, rng_call;
    long long max_iters;
    ReportDev* rep;
};

// ------------------------------------------------------------------------------------------------
// small numeric helpers (NumPy-compatible promotion: C++ usual arithmetic conversions on
// float/double are the same lattice as NumPy's for float32/float64 arrays; Python scalars are
// "weak", so literals below are always cast to the array type first)
// ------------------------------------------------------------------------------------------------
template <class A, class B>
using prom_t = decltype(A() + B());

template <class T>
__device__ __forceinline__ T ldg(const T* p) {
    return __ldg(p);
}

// np.deg2rad: x * (pi/180) evaluated in the array dtype (npy_deg2rad / npy_deg2radf)
__device__ __forceinline__ float deg2rad_np(float x) { return x * (float)(3.14159265358979323846 / 180.0); }
__device__ __forceinline__ double deg2rad_np(double x) { return x * (3.14159265358979323846 / 180.0); }
__device__ __forceinline__ float cos_np(float x) { return cosf(x); }
__device__ __forceinline__ double cos_np(double x) { return cos(x); }

// A value with NumPy's dtype tag: interpolated velocities are float32 only when grid coordinates,
// field data and the sampled position are all float32 (stage 1) -- the tag decides in which
// precision the in-place spherical division is rounded (_xinterpolators.py:182-184).
struct Val {
    double v;
    bool f32;
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (Wiener increments of DiffusionUniformKh)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void wiener_normals(unsigned long long seed, unsigned long long rng_call, long long iter,
                                               long long pid, double& zx, double& zy) {
    uint32_t r[4];
    // counter = (pid_lo, pid_hi, iteration, call index); key = seed
    philox4x32_10((uint32_t)pid, (uint32_t)((unsigned long long)pid >> 32), (uint32_t)iter, (uint32_t)rng_call,
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double two_m32 = 2.3283064365386963e-10;  // 2^-32
    double u1 = ((double)r[0] + 0.5) * two_m32;
    double u2 = ((double)r[1] + 0.5) * two_m32;
    double rad = sqrt(-2.0 * log(u1));
    double ang = 6.283185307179586476925 * u2;
    zx = rad * cos(ang);
    zy = rad * sin(ang);
}

// ------------------------------------------------------------------------------------------------
// 1-D axis search with a register-resident cell (reference _core/index_search.py:20-62)
//   idx = clip(searchsorted(arr, x, 'left') - 1, 0, n-2);  b = (x - arr[idx]) / (arr[idx+1] - arr[idx])
//   idx = -2 if x < arr[0];  idx = -1 if x > arr[-1]
// The previous cell [lo, hi] is kept in registers: lo < x <= hi is exactly the searchsorted
// condition for that cell, so a hit costs no memory access.
// ------------------------------------------------------------------------------------------------
template <class A>
struct AxisCell {
    int idx;  // raw result of the last search (may be a negative sentinel)
    A lo, hi;
};

template <class P, class A>
__device__ __forceinline__ prom_t<P, A> axis_search(const A* __restrict__ arr, int n, P x, AxisCell<A>& c) {
    using R = prom_t<P, A>;
    if (n < 2) {  // index_search.py:45-46
        c.idx = 0;
        return (R)0;
    }
    const R xr = (R)x;
    if (!(c.idx >= 0 && xr > (R)c.lo && xr <= (R)c.hi)) {
        int l = 0, h = n;  // first i with arr[i] >= x   (side="left")
        while (l < h) {
            int m = (l + h) >> 1;
            if ((R)ldg(arr + m) < xr) l = m + 1; else h = m;
        }
        if (x != x) l = n;  // NaN sorts last
        int i = min(max(l - 1, 0), n - 2);
        c.lo = ldg(arr + i);
        c.hi = ldg(arr + i + 1);
        c.idx = i;
        if (xr < (R)ldg(arr)) c.idx = -2;          // LEFT_OUT_OF_BOUNDS
        if (xr > (R)ldg(arr + n - 1)) c.idx = -1;  // RIGHT_OUT_OF_BOUNDS
    }
    return (xr - (R)c.lo) / (R)(c.hi - c.lo);  // denominator rounded in the coordinate dtype
}

// ------------------------------------------------------------------------------------------------
// corner cache + XLinear (reference interpolators/_xinterpolators.py:78-153)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long wrap_idx(int i, int n) {  // NumPy negative-index wrap
    int w = i < 0 ? i + n : i;
    return (long long)min(max(w, 0), n - 1);
}
__device__ __forceinline__ long long up_idx(int i, int n) {  // np.clip(i + 1, 0, n - 1)
    return (long long)min(max(i + 1, 0), n - 1);
}

template <class D, int NC>
struct Corners {
    int ti, zi, yi, xi;  // key of the block held in v (INT_MIN: empty)
    D v[NC][16];         // [component][(t*2+z)*4 + y*2 + x]

    __device__ __forceinline__ void fill(const FieldDev& f, int nti, int nzi, int nyi, int nxi) {
        ti = nti; zi = nzi; yi = nyi; xi = nxi;
        long long ot[2] = {wrap_idx(nti, f.T) * f.sT, up_idx(nti, f.T) * f.sT};
        long long oz[2] = {wrap_idx(nzi, f.Z) * f.sZ, up_idx(nzi, f.Z) * f.sZ};
        long long oy[2] = {wrap_idx(nyi, f.Y) * f.sY, up_idx(nyi, f.Y) * f.sY};
        long long ox[2] = {wrap_idx(nxi, f.X) * f.sX, up_idx(nxi, f.X) * f.sX};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const D* __restrict__ base = (const D*)f.p[c];
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[c][k] = ldg(base + ot[k >> 3] + oz[(k >> 2) & 1] + oy[(k >> 1) & 1] + ox[k & 1]);
        }
    }
};

// bilinear in (eta, xsi) on 4 values of type C:  _xinterpolators.py:147-152, evaluated left to right
template <class C, class TY, class TX>
__device__ __forceinline__ Val bilinear(const C (&c)[4], TY eta, TX xsi) {
    auto r = (1 - xsi) * (1 - eta) * c[0] + xsi * (1 - eta) * c[1] + (1 - xsi) * eta * c[2] + xsi * eta * c[3];
    return Val{(double)r, std::is_same<decltype(r), float>::value};
}

// Z-lerp (only when zeta > 0: `lenZ`, _xinterpolators.py:131,141-145) then bilinear
template <class C, class TZ, class TY, class TX>
__device__ __forceinline__ Val zlerp_bilinear(const C (&c)[8], TZ zeta, TY eta, TX xsi) {
    if (zeta > 0) {
        using R = prom_t<C, TZ>;
        R r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = c[k] * (1 - zeta) + c[4 + k] * zeta;
        return bilinear<R, TY, TX>(r, eta, xsi);
    } else {
        C r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = c[k];
        return bilinear<C, TY, TX>(r, eta, xsi);
    }
}

// T-lerp (only when tau > 0: `lenT`, _xinterpolators.py:130,135-139), then Z, then YX.
// The reference decides lenT/lenZ per batch (any(tau > 0)); per particle the arithmetic is the
// same, and so is the dtype whenever the particles of a batch share their clock.
template <class D, class TT, class TZ, class TY, class TX>
__device__ __forceinline__ Val xlinear(const D (&v)[16], TT tau, TZ zeta, TY eta, TX xsi) {
    if (tau > 0) {
        using R = prom_t<D, TT>;
        R r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = v[k] * (1 - tau) + v[8 + k] * tau;
        return zlerp_bilinear<R, TZ, TY, TX>(r, zeta, eta, xsi);
    } else {
        D r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = v[k];
        return zlerp_bilinear<D, TZ, TY, TX>(r, zeta, eta, xsi);
    }
}

// ------------------------------------------------------------------------------------------------
// per-particle evaluation state
// ------------------------------------------------------------------------------------------------
template <class A, class D, int NC>
struct EvalCtx {
    AxisCell<A> cx, cy, cz;
    AxisCell<double> ct;
    Corners<D, NC> cor;
    int state;
    int ei;
    unsigned int refills;
    bool out_of_time;
};

// VectorField.eval for one particle (reference _core/field.py:250-304,307-405 with
// XLinear_Velocity, _xinterpolators.py:169-190).  PZ/PY/PX: dtype of the sampled position
// (float32 = the particle's own arrays, float64 = an RK stage position).
template <class A, class D, bool HAS_TIME, int NC, class PZ, class PY, class PX>
__device__ __forceinline__ void eval_uvw(const GridDev& g, const FieldDev& f, EvalCtx<A, D, NC>& e, double t, PZ z, PY y,
                                         PX x, Val& u, Val& v, Val& w) {
    using TT = typename std::conditional<HAS_TIME, double, float>::type;
    using TZ = prom_t<PZ, A>;
    using TY = prom_t<PY, A>;
    using TX = prom_t<PX, A>;

    // -- time index (index_search.py:65-91).  Outside [0, length]: the reference raises
    //    OutsideTimeInterval, the kernel's view gets state 70 and the sample is (0, 0, 0).
    TT tau = 0;
    int ti = 0;
    if (HAS_TIME) {
        if (!(0 <= t && t <= g.time_len)) {
            e.state = PB_ERROR_OUTSIDE_TIME_INTERVAL;
            e.out_of_time = true;
            u = Val{0.0, false}; v = u; w = u;
            return;
        }
        tau = (TT)axis_search<double, double>(g.time, g.nt, t, e.ct);
        ti = e.ct.idx;
    }
    // -- XGrid.search (xgrid.py:316-356)
    TZ zeta = 0;
    int zi = 0;
    if (g.nz > 0) {
        zeta = axis_search<PZ, A>((const A*)g.depth, g.nz, z, e.cz);
        zi = e.cz.idx;
    }
    TY eta = axis_search<PY, A>((const A*)g.lat, g.ny, y, e.cy);
    TX xsi = axis_search<PX, A>((const A*)g.lon, g.nx, x, e.cx);
    const int yi = e.cy.idx, xi = e.cx.idx;

    // -- particles.ei[:, igrid] = ravel_index (field.py:307-317, basegrid.py:259-278); int64 -> int32
    long long r = (long long)yi * g.xdim + (long long)xi;
    if (g.nz > 0) r += (long long)zi * (g.ydim * g.xdim);
    e.ei = (int)r;

    // -- state from positions (field.py:327-356).  X/Y index -2 is NOT an error in the reference.
    int s = e.state;
    if (xi == -1 || yi == -1 || zi == -1) s = max(s, (int)PB_ERROR_OUT_OF_BOUNDS);
    if (zi == -2) s = max(s, (int)PB_ERROR_THROUGH_SURFACE);

    // -- corner block: gather from HBM only when the bracketing block changed
    if (e.cor.ti != ti || e.cor.zi != zi || e.cor.yi != yi || e.cor.xi != xi) {
        e.cor.fill(f, ti, zi, yi, xi);
        e.refills++;
    }

    u = xlinear<D, TT, TZ, TY, TX>(e.cor.v[0], tau, zeta, eta, xsi);
    v = xlinear<D, TT, TZ, TY, TX>(e.cor.v[1], tau, zeta, eta, xsi);
    if (g.spherical) {  // u /= deg2m * cos(deg2rad(y)); v /= deg2m   (in-place: result keeps u's dtype)
        PY conv = (PY)g.deg2m * cos_np(deg2rad_np(y));
        if (u.f32 && std::is_same<PY, float>::value) {
            u.v = (double)((float)u.v / (float)conv);
        } else {
            double q = u.v / (double)conv;
            u.v = u.f32 ? (double)(float)q : q;
        }
        v.v = v.f32 ? (double)((float)v.v / (float)g.deg2m) : v.v / g.deg2m;
    }
    if (NC == 3) {
        w = xlinear<D, TT, TZ, TY, TX>(e.cor.v[NC - 1], tau, zeta, eta, xsi);
    } else {
        w = Val{0.0, u.f32};
    }
    // -- NaN -> ErrorInterpolation, then out-of-bounds samples -> 0 (field.py:288-290,359-378)
    if (u.v != u.v || v.v != v.v || w.v != w.v) s = max(s, (int)PB_ERROR_INTERPOLATION);
    if (xi < 0 || yi < 0 || zi < 0) {
        u.v = 0.0; v.v = 0.0; w.v = 0.0;
    }
    e.state = s;
}

// u * 0.5 keeps u's dtype (Python float is weak)
__device__ __forceinline__ double half_of(const Val& a) { return a.f32 ? (double)((float)a.v * 0.5f) : a.v * 0.5; }

// ------------------------------------------------------------------------------------------------
// the kernel: Kernel.execute's loop, one lane per particle
// ------------------------------------------------------------------------------------------------
template <class A, class D, bool HAS_TIME, int NC>
__global__ void __launch_bounds__(128) advect_kernel(const AdvectParams p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long my_steps = 0, my_refills = 0;
    int final_state = 0;
    long long my_iters = 0;
    bool errored = false, deleted = false, oot = false;
    long long err_iter = LLONG_MAX;

    if (i < p.P.n) {
        float x = p.P.x[i], y = p.P.y[i], z = p.P.z[i];
        float dx = p.P.dx[i], dy = p.P.dy[i], dz = p.P.dz[i];
        double t = p.P.t[i];
        const long long pid = p.diffusion ? p.P.pid[i] : 0;

        EvalCtx<A, D, NC> e;
        e.cx.idx = e.cy.idx = e.cz.idx = e.ct.idx = -100;
        e.cx.lo = e.cx.hi = e.cy.lo = e.cy.hi = e.cz.lo = e.cz.hi = (A)0;
        e.ct.lo = e.ct.hi = 0.0;
        e.cor.ti = e.cor.zi = e.cor.yi = e.cor.xi = INT_MIN;
        e.state = PB_EVALUATE;  // kernel.py:188
        e.ei = p.P.ei[i];
        e.refills = 0;
        e.out_of_time = false;

        const int sign = p.dt > 0 ? 1 : -1;
        const bool three_d = (p.scheme == PB_ADVECTION_RK4_3D || p.scheme == PB_ADVECTION_RK2_3D);
        const int nstage = (p.scheme == PB_ADVECTION_EE) ? 1 : ((p.scheme == PB_ADVECTION_RK2 || p.scheme == PB_ADVECTION_RK2_3D) ? 2 : 4);

        long long it = 0;
        for (;; ++it) {
            if (p.max_iters >= 0 && it >= p.max_iters) break;
            const double tte = sign * (p.endtime - t);                       // kernel.py:191
            if (!((e.state == PB_SUCCESS || e.state == PB_EVALUATE) && tte >= 0)) break;  // :193-195
            // adapt dt to end exactly on endtime (:199-203)
            const double dtp = (sign == 1) ? fmax(fmin(p.dt, tte), 0.0) : fmin(fmax(p.dt, -tte), 0.0);
            my_steps++;

            // ---- advection kernel (kernels/_advection.py) ----
            Val u1, v1, w1, uk, vk, wk;
            eval_uvw<A, D, HAS_TIME, NC, float, float, float>(p.g, p.f, e, t, z, y, x, u1, v1, w1);
            double su = u1.v, sv = v1.v, sw = w1.v;  // running RK4 sums, left to right
            uk = u1; vk = v1; wk = w1;
            for (int k = 1; k < nstage; ++k) {
                // stage position: x + u*0.5*dt (k = 1, 2) or x + u*dt (k = 3)
                const bool full = (k == 3);
                const double xs = (double)x + (full ? uk.v : half_of(uk)) * dtp;
                const double ys = (double)y + (full ? vk.v : half_of(vk)) * dtp;
                const double ts = t + (full ? dtp : 0.5 * dtp);
                if (three_d) {
                    const double zs = (double)z + (full ? wk.v : half_of(wk)) * dtp;
                    eval_uvw<A, D, HAS_TIME, NC, double, double, double>(p.g, p.f, e, ts, zs, ys, xs, uk, vk, wk);
                } else {
                    eval_uvw<A, D, HAS_TIME, NC, float, double, double>(p.g, p.f, e, ts, z, ys, xs, uk, vk, wk);
                }
                if (nstage == 4) {
                    const double m = (k == 3) ? 1.0 : 2.0;  // u1 + 2*u2 + 2*u3 + u4
                    su = su + m * uk.v; sv = sv + m * vk.v; sw = sw + m * wk.v;
                }
            }
            double ddx, ddy, ddz;
            if (nstage == 4) {
                ddx = su / 6.0 * dtp; ddy = sv / 6.0 * dtp; ddz = sw / 6.0 * dtp;
            } else {  // EE: u1*dt ; RK2: u2*dt
                ddx = uk.v * dtp; ddy = vk.v * dtp; ddz = wk.v * dtp;
            }
            dx = (float)((double)dx + ddx);
            dy = (float)((double)dy + ddy);
            if (three_d) dz = (float)((double)dz + ddz);

            // ---- DiffusionUniformKh (kernels/_advectiondiffusion.py:120-153) ----
            if (p.diffusion) {
                double zx, zy;
                wiener_normals(p.seed, p.rng_call, it, pid, zx, zy);
                const double sq = sqrt(fabs(dtp));
                const double dWx = zx * sq, dWy = zy * sq;
                double khz = p.kh_zonal, khm = p.kh_meridional;
                if (p.kh_spherical) {
                    const float ang = (y * (float)3.14159265358979323846) / 180.0f;  // lat * np.pi / 180 in f32
                    const float m = (float)p.kh_deg2m * cosf(ang);
                    khz = khz / (double)(m * m);
                    khm = khm / (p.kh_deg2m * p.kh_deg2m);
                }
                const double bx = sqrt(2 * khz), by = sqrt(2 * khm);
                dx = (float)((double)dx + bx * dWx);
                dy = (float)((double)dy + by * dWy);
                e.ei = 0;  // the constant-field evals overwrite ei[:, -1] with cell 0 (model.py:292-318)
            }
            // ---- trailing error handler: every error state becomes Delete ----
            if (p.delete_on_error && e.state >= 50) e.state = PB_DELETE;

            // ---- position update only for particles still in a normal state (kernel.py:108-116,220-222)
            if (e.state == PB_EVALUATE || e.state == PB_SUCCESS) {
                x = x + dx; y = y + dy; z = z + dz;
                t = t + dtp;
                dx = 0.f; dy = 0.f; dz = 0.f;
            }
            if (e.state == PB_EVALUATE && t == p.endtime) e.state = PB_END_OF_LOOP;  // :229-230
            if (e.state == PB_DELETE) { deleted = true; ++it; break; }
            if (e.state >= 50) { errored = true; err_iter = it; ++it; break; }
        }
        my_iters = it;
        my_refills = e.refills;
        oot = e.out_of_time;
        final_state = e.state;
        p.P.x[i] = x; p.P.y[i] = y; p.P.z[i] = z;
        p.P.dx[i] = dx; p.P.dy[i] = dy; p.P.dz[i] = dz;
        p.P.t[i] = t;
        p.P.state[i] = e.state;
        p.P.ei[i] = e.ei;
    }

    // ---- report: warp-reduce then one atomic per warp ----
    const unsigned full = 0xffffffffu;
    unsigned long long s_steps = my_steps, s_ref = my_refills;
    unsigned n_err = errored, n_del = deleted, n_oot = oot;
    long long mx_it = my_iters, mn_err = err_iter;
    int mx_state = final_state;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_steps += __shfl_xor_sync(full, s_steps, o);
        s_ref += __shfl_xor_sync(full, s_ref, o);
        n_err += __shfl_xor_sync(full, n_err, o);
        n_del += __shfl_xor_sync(full, n_del, o);
        n_oot += __shfl_xor_sync(full, n_oot, o);
        mx_it = max(mx_it, __shfl_xor_sync(full, mx_it, o));
        mn_err = min(mn_err, __shfl_xor_sync(full, mn_err, o));
        mx_state = max(mx_state, __shfl_xor_sync(full, mx_state, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (s_steps) atomicAdd(&p.rep->particle_steps, s_steps);
        if (s_ref) atomicAdd(&p.rep->cache_refills, s_ref);
        if (n_err) atomicAdd(&p.rep->n_error, (unsigned long long)n_err);
        if (n_del) atomicAdd(&p.rep->n_deleted, (unsigned long long)n_del);
        if (n_oot) atomicAdd(&p.rep->n_out_of_time, (unsigned long long)n_oot);
        if (mx_it) atomicMax(&p.rep->max_iters_done, mx_it);
        if (mn_err != LLONG_MAX) atomicMin(&p.rep->first_error_iter, mn_err);
        if (mx_state) atomicMax(&p.rep->max_state, mx_state);
    }
}

// whole-view OutsideTimeInterval flag (index_search.py:85-86 + field.py:31-44)
__global__ void flag_view_kernel(ParticlesDev P, double dt, double endtime) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const int sign = dt > 0 ? 1 : -1;
    const int s = P.state[i];
    const double tte = sign * (endtime - P.t[i]);
    if ((s == PB_SUCCESS || s == PB_EVALUATE) && tte >= 0) P.state[i] = PB_ERROR_OUTSIDE_TIME_INTERVAL;
}

__global__ void normals_kernel(unsigned long long seed, unsigned long long rng_call, long long iter, long long n,
                               const long long* pid, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double zx, zy;
    wiener_normals(seed, rng_call, iter, pid[i], zx, zy);
    out[2 * i] = zx;
    out[2 * i + 1] = zy;
}

// ------------------------------------------------------------------------------------------------
// host-side engine
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int32_t ensure(size_t n) {
        if (n <= bytes) return PB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        CK(cudaMalloc(&p, n));
        bytes = n;
        return PB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

struct pb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
    // grid
    DevBuf lon, lat, depth, time;
    GridDev g{};
    bool have_grid = false;
    int coord_f64 = 0;
    // fields
    DevBuf fbuf[3];
    const void* fptr[3] = {nullptr, nullptr, nullptr};
    int f_f64[3] = {0, 0, 0};
    long long fshape[3][4] = {};
    // particles
    DevBuf px, py, pz, pdx, pdy, pdz, pt, pstate, pei, ppid;
    DevBuf snap;  // snapshot of all particle arrays
    long long n = 0;
    bool have_pid = false;
    // report
    ReportDev* d_rep = nullptr;
    ReportDev* h_rep = nullptr;  // pinned
    pb_report last{};
    bool pending = false;
};

static void zero_report(ReportDev& r) {
    memset(&r, 0, sizeof(r));
    r.first_error_iter = LLONG_MAX;
}

template <class A, class D, bool HT, int NC>
static cudaError_t launch(const AdvectParams& p, cudaStream_t s) {
    const int block = 128;
    const long long grid = (p.P.n + block - 1) / block;
    advect_kernel<A, D, HT, NC><<<(unsigned)grid, block, 0, s>>>(p);
    return cudaGetLastError();
}

template <class A, class D>
static cudaError_t launch_ad(const AdvectParams& p, bool ht, int nc, cudaStream_t s) {
    if (ht) return nc == 3 ? launch<A, D, true, 3>(p, s) : launch<A, D, true, 2>(p, s);
    return nc == 3 ? launch<A, D, false, 3>(p, s) : launch<A, D, false, 2>(p, s);
}

extern "C" {

int32_t pb_abi_version(void) { return PB_ABI_VERSION; }
const char* pb_last_error_string(void) { return g_err.c_str(); }

int32_t pb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t pb_engine_create(int32_t device, pb_engine** out) {
    if (!out) return fail(PB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = pb_device_count();
    if (n <= 0) return fail(PB_ERR_NO_DEVICE, "no CUDA device visible: parcels_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(PB_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(PB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    pb_engine* e = new pb_engine();
    e->device = device;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
    CK(cudaEventCreate(&e->tev0));
    CK(cudaEventCreate(&e->tev1));
    CK(cudaMalloc(&e->d_rep, sizeof(ReportDev)));
    CK(cudaMallocHost(&e->h_rep, sizeof(ReportDev)));
    *out = e;
    return PB_OK;
}

void pb_engine_destroy(pb_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (DevBuf* b : {&e->lon, &e->lat, &e->depth, &e->time, &e->fbuf[0], &e->fbuf[1], &e->fbuf[2], &e->px, &e->py, &e->pz,
                      &e->pdx, &e->pdy, &e->pdz, &e->pt, &e->pstate, &e->pei, &e->ppid, &e->snap})
        b->release();
    if (e->d_rep) cudaFree(e->d_rep);
    if (e->h_rep) cudaFreeHost(e->h_rep);
    cudaEventDestroy(e->ev0);
    cudaEventDestroy(e->ev1);
    cudaEventDestroy(e->tev0);
    cudaEventDestroy(e->tev1);
    cudaStreamDestroy(e->stream);
    delete e;
}

int32_t pb_engine_synchronize(pb_engine* e) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return PB_OK;
}

int32_t pb_timer_begin(pb_engine* e) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    CK(cudaSetDevice(e->device));
    CK(cudaEventRecord(e->tev0, e->stream));
    return PB_OK;
}

int32_t pb_timer_end_ms(pb_engine* e, float* ms) {
    if (!e || !ms) return fail(PB_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(e->device));
    CK(cudaEventRecord(e->tev1, e->stream));
    CK(cudaEventSynchronize(e->tev1));
    CK(cudaEventElapsedTime(ms, e->tev0, e->tev1));
    return PB_OK;
}

static int32_t upload(pb_engine* e, DevBuf& b, const void* src, size_t bytes) {
    int32_t rc = b.ensure(bytes ? bytes : 1);
    if (rc) return rc;
    if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream));
    return PB_OK;
}

int32_t pb_grid_upload_rectilinear(pb_engine* e, const void* lon, int64_t nx, const void* lat, int64_t ny,
                                   const void* depth, int64_t nz, int32_t coord_is_f64, const double* time_s,
                                   int64_t nt, int32_t spherical, double deg2m, int64_t xdim_cells,
                                   int64_t ydim_cells, int64_t zdim_cells) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    if (!lon || !lat || nx < 2 || ny < 2)
        return fail(PB_ERR_INVALID, "rectilinear grid needs lon and lat with >= 2 nodes (got nx=%lld ny=%lld)", (long long)nx, (long long)ny);
    if (nx > INT_MAX || ny > INT_MAX || nz > INT_MAX || nt > INT_MAX) return fail(PB_ERR_INVALID, "axis too long");
    CK(cudaSetDevice(e->device));
    const size_t es = coord_is_f64 ? 8 : 4;
    int32_t rc;
    if ((rc = upload(e, e->lon, lon, nx * es))) return rc;
    if ((rc = upload(e, e->lat, lat, ny * es))) return rc;
    if (!depth) nz = 0;
    if ((rc = upload(e, e->depth, depth, nz * es))) return rc;
    std::vector<double> tnorm;
    if (!time_s || nt < 2) nt = 0;
    if (nt) {  // seconds since the interval start (index_search.py:88)
        tnorm.resize(nt);
        for (int64_t k = 0; k < nt; ++k) tnorm[k] = time_s[k] - time_s[0];
        for (int64_t k = 1; k < nt; ++k)
            if (!(tnorm[k] > tnorm[k - 1])) return fail(PB_ERR_INVALID, "time axis must be strictly increasing");
    }
    if ((rc = upload(e, e->time, tnorm.data(), nt * sizeof(double)))) return rc;
    CK(cudaStreamSynchronize(e->stream));  // tnorm goes out of scope
    GridDev& g = e->g;
    g.lon = e->lon.p; g.lat = e->lat.p; g.depth = nz ? e->depth.p : nullptr;
    g.time = nt ? (const double*)e->time.p : nullptr;
    g.nx = (int)nx; g.ny = (int)ny; g.nz = (int)nz; g.nt = (int)nt;
    g.spherical = spherical ? 1 : 0;
    g.deg2m = spherical ? deg2m : 1.0;
    g.time_len = nt ? tnorm[nt - 1] : 0.0;
    g.xdim = xdim_cells; g.ydim = ydim_cells; g.zdim = zdim_cells;
    e->coord_f64 = coord_is_f64 ? 1 : 0;
    e->have_grid = true;
    return PB_OK;
}

static int32_t set_field(pb_engine* e, int32_t slot, const void* dev, int32_t is_f64, int64_t T, int64_t Z, int64_t Y, int64_t X) {
    e->fptr[slot] = dev;
    e->f_f64[slot] = is_f64 ? 1 : 0;
    e->fshape[slot][0] = T; e->fshape[slot][1] = Z; e->fshape[slot][2] = Y; e->fshape[slot][3] = X;
    return PB_OK;
}

int32_t pb_field_upload(pb_engine* e, int32_t slot, const void* data, int32_t data_is_f64, int64_t T, int64_t Z,
                        int64_t Y, int64_t X) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    if (slot < 0 || slot > 2) return fail(PB_ERR_INVALID, "bad field slot %d", slot);
    if (!data || T < 1 || Z < 1 || Y < 1 || X < 1) return fail(PB_ERR_INVALID, "bad field shape");
    if (T > INT_MAX || Z > INT_MAX || Y > INT_MAX || X > INT_MAX) return fail(PB_ERR_INVALID, "field dim too long");
    CK(cudaSetDevice(e->device));
    const size_t bytes = (size_t)T * Z * Y * X * (data_is_f64 ? 8 : 4);
    int32_t rc = upload(e, e->fbuf[slot], data, bytes);
    if (rc) return rc;
    CK(cudaStreamSynchronize(e->stream));
    return set_field(e, slot, e->fbuf[slot].p, data_is_f64, T, Z, Y, X);
}

int32_t pb_field_attach_device(pb_engine* e, int32_t slot, const void* dev_data, int32_t data_is_f64, int64_t T,
                               int64_t Z, int64_t Y, int64_t X) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    if (slot < 0 || slot > 2) return fail(PB_ERR_INVALID, "bad field slot %d", slot);
    if (!dev_data || T < 1 || Z < 1 || Y < 1 || X < 1) return fail(PB_ERR_INVALID, "bad field shape");
    if (T > INT_MAX || Z > INT_MAX || Y > INT_MAX || X > INT_MAX) return fail(PB_ERR_INVALID, "field dim too long");
    e->fbuf[slot].release();
    return set_field(e, slot, dev_data, data_is_f64, T, Z, Y, X);
}

int32_t pb_field_clear(pb_engine* e, int32_t slot) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    if (slot < 0 || slot > 2) return fail(PB_ERR_INVALID, "bad field slot %d", slot);
    e->fbuf[slot].release();
    e->fptr[slot] = nullptr;
    return PB_OK;
}

int32_t pb_particles_upload(pb_engine* e, int64_t n, const float* x, const float* y, const float* z, const float* dx,
                            const float* dy, const float* dz, const double* t, const int32_t* state, const int32_t* ei,
                            const int64_t* particle_id) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    if (n < 0) return fail(PB_ERR_INVALID, "n < 0");
    if (n && (!x || !y || !z || !t || !state || !ei)) return fail(PB_ERR_INVALID, "NULL particle array");
    CK(cudaSetDevice(e->device));
    int32_t rc;
    if ((rc = upload(e, e->px, x, n * 4))) return rc;
    if ((rc = upload(e, e->py, y, n * 4))) return rc;
    if ((rc = upload(e, e->pz, z, n * 4))) return rc;
    for (auto pr : {std::make_pair(&e->pdx, dx), std::make_pair(&e->pdy, dy), std::make_pair(&e->pdz, dz)}) {
        if ((rc = pr.first->ensure(n ? n * 4 : 1))) return rc;
        if (pr.second) { if (n) CK(cudaMemcpyAsync(pr.first->p, pr.second, n * 4, cudaMemcpyHostToDevice, e->stream)); }
        else if (n) CK(cudaMemsetAsync(pr.first->p, 0, n * 4, e->stream));
    }
    if ((rc = upload(e, e->pt, t, n * 8))) return rc;
    if ((rc = upload(e, e->pstate, state, n * 4))) return rc;
    if ((rc = upload(e, e->pei, ei, n * 4))) return rc;
    e->have_pid = particle_id != nullptr;
    if (particle_id && (rc = upload(e, e->ppid, particle_id, n * 8))) return rc;
    e->n = n;
    CK(cudaStreamSynchronize(e->stream));
    return PB_OK;
}

int32_t pb_particles_download(pb_engine* e, int64_t n, float* x, float* y, float* z, float* dx, float* dy, float* dz,
                              double* t, int32_t* state, int32_t* ei) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    if (n != e->n) return fail(PB_ERR_INVALID, "download of %lld particles but %lld are resident", (long long)n, (long long)e->n);
    CK(cudaSetDevice(e->device));
    struct { void* dst; void* src; size_t es; } cp[] = {
        {x, e->px.p, 4}, {y, e->py.p, 4}, {z, e->pz.p, 4}, {dx, e->pdx.p, 4}, {dy, e->pdy.p, 4}, {dz, e->pdz.p, 4},
        {t, e->pt.p, 8}, {state, e->pstate.p, 4}, {ei, e->pei.p, 4}};
    for (auto& c : cp)
        if (c.dst && n) CK(cudaMemcpyAsync(c.dst, c.src, n * c.es, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return PB_OK;
}

static const size_t kSnapOff[10] = {0, 4, 8, 12, 16, 20, 24, 32, 36, 40};  // per-particle byte offsets x..ei (t at 24)

int32_t pb_particles_snapshot(pb_engine* e) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    CK(cudaSetDevice(e->device));
    const size_t n = (size_t)e->n;
    int32_t rc = e->snap.ensure(n ? n * 40 : 1);
    if (rc) return rc;
    char* s = (char*)e->snap.p;
    struct { void* src; size_t es; } a[] = {{e->px.p, 4}, {e->py.p, 4}, {e->pz.p, 4}, {e->pdx.p, 4}, {e->pdy.p, 4},
                                            {e->pdz.p, 4}, {e->pt.p, 8}, {e->pstate.p, 4}, {e->pei.p, 4}};
    size_t off = 0;
    for (auto& c : a) {
        if (n) CK(cudaMemcpyAsync(s + off, c.src, n * c.es, cudaMemcpyDeviceToDevice, e->stream));
        off += n * c.es;
    }
    (void)kSnapOff;
    return PB_OK;
}

int32_t pb_particles_restore(pb_engine* e) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    CK(cudaSetDevice(e->device));
    const size_t n = (size_t)e->n;
    if (e->snap.bytes < n * 40) return fail(PB_ERR_STATE, "no snapshot to restore");
    char* s = (char*)e->snap.p;
    struct { void* dst; size_t es; } a[] = {{e->px.p, 4}, {e->py.p, 4}, {e->pz.p, 4}, {e->pdx.p, 4}, {e->pdy.p, 4},
                                            {e->pdz.p, 4}, {e->pt.p, 8}, {e->pstate.p, 4}, {e->pei.p, 4}};
    size_t off = 0;
    for (auto& c : a) {
        if (n) CK(cudaMemcpyAsync(c.dst, s + off, n * c.es, cudaMemcpyDeviceToDevice, e->stream));
        off += n * c.es;
    }
    return PB_OK;
}

int64_t pb_particles_count(pb_engine* e) { return e ? e->n : -1; }

int32_t pb_advect_async(pb_engine* e, const pb_advect_args* a) {
    if (!e || !a) return fail(PB_ERR_INVALID, "NULL argument");
    if (!e->have_grid) return fail(PB_ERR_STATE, "pb_advect before pb_grid_upload_*");
    if (!e->fptr[0] || !e->fptr[1]) return fail(PB_ERR_STATE, "pb_advect before U and V were uploaded");
    if (a->dt == 0.0 || a->dt != a->dt) return fail(PB_ERR_INVALID, "dt must be a non-zero number");
    int nc;
    switch (a->scheme) {
        case PB_ADVECTION_EE: case PB_ADVECTION_RK2: case PB_ADVECTION_RK4: nc = 2; break;
        case PB_ADVECTION_RK2_3D: case PB_ADVECTION_RK4_3D: nc = 3; break;
        default: return fail(PB_ERR_INVALID, "unknown scheme %d", a->scheme);
    }
    if (nc == 3 && !e->fptr[2]) return fail(PB_ERR_STATE, "3-D scheme needs a W field (fieldset.UVW)");
    for (int c = 1; c < nc; ++c) {
        if (e->f_f64[c] != e->f_f64[0]) return fail(PB_ERR_INVALID, "U, V, W must share one dtype");
        for (int d = 0; d < 4; ++d)
            if (e->fshape[c][d] != e->fshape[0][d]) return fail(PB_ERR_INVALID, "U, V, W must share one shape on an A-grid");
    }
    if (a->diffusion && !e->have_pid) return fail(PB_ERR_STATE, "diffusion needs particle_id (RNG counter)");
    const long long T = e->fshape[0][0], Z = e->fshape[0][1], Y = e->fshape[0][2], X = e->fshape[0][3];
    // A-grid data must match the node counts wherever the dim is indexed
    if ((X > 1 && X != e->g.nx) || (Y > 1 && Y != e->g.ny) || (Z > 1 && e->g.nz > 0 && Z != e->g.nz) ||
        (T > 1 && e->g.nt > 0 && T != e->g.nt))
        return fail(PB_ERR_INVALID, "field shape (%lld,%lld,%lld,%lld) does not match grid nodes (nt=%d nz=%d ny=%d nx=%d)", T, Z, Y, X,
                    e->g.nt, e->g.nz, e->g.ny, e->g.nx);
    if ((T > 1) != (e->g.nt > 0)) return fail(PB_ERR_INVALID, "field has %lld time levels but the grid time axis has %d", T, e->g.nt);
    if (Z > 1 && e->g.nz == 0) return fail(PB_ERR_INVALID, "field has a depth dimension but the grid has no Z axis");
    CK(cudaSetDevice(e->device));

    AdvectParams p{};
    p.g = e->g;
    for (int c = 0; c < 3; ++c) p.f.p[c] = e->fptr[c];
    p.f.T = (int)T; p.f.Z = (int)Z; p.f.Y = (int)Y; p.f.X = (int)X;
    p.f.sX = X > 1 ? 1 : 0;
    p.f.sY = Y > 1 ? X : 0;
    p.f.sZ = Z > 1 ? X * Y : 0;
    p.f.sT = T > 1 ? X * Y * Z : 0;
    p.P = ParticlesDev{(float*)e->px.p, (float*)e->py.p, (float*)e->pz.p, (float*)e->pdx.p, (float*)e->pdy.p,
                       (float*)e->pdz.p, (double*)e->pt.p, (int*)e->pstate.p, (int*)e->pei.p, (long long*)e->ppid.p, e->n};
    p.scheme = a->scheme; p.diffusion = a->diffusion; p.delete_on_error = a->delete_on_error;
    p.kh_spherical = a->kh_spherical;
    p.dt = a->dt; p.endtime = a->endtime;
    p.kh_zonal = a->kh_zonal; p.kh_meridional = a->kh_meridional; p.kh_deg2m = a->kh_deg2m;
    p.seed = a->seed; p.rng_call = a->rng_call; p.max_iters = a->max_iters;
    p.rep = e->d_rep;

    zero_report(*e->h_rep);
    CK(cudaMemcpyAsync(e->d_rep, e->h_rep, sizeof(ReportDev), cudaMemcpyHostToDevice, e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    if (e->n > 0) {
        const bool ht = e->g.nt > 0;
        cudaError_t ce;
        if (e->coord_f64) ce = e->f_f64[0] ? launch_ad<double, double>(p, ht, nc, e->stream) : launch_ad<double, float>(p, ht, nc, e->stream);
        else ce = e->f_f64[0] ? launch_ad<float, double>(p, ht, nc, e->stream) : launch_ad<float, float>(p, ht, nc, e->stream);
        if (ce != cudaSuccess) return fail(PB_ERR_CUDA, "advect_kernel launch failed: %s", cudaGetErrorString(ce));
    }
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaMemcpyAsync(e->h_rep, e->d_rep, sizeof(ReportDev), cudaMemcpyDeviceToHost, e->stream));
    e->pending = true;
    return PB_OK;
}

int32_t pb_last_report(pb_engine* e, pb_report* rep) {
    if (!e || !rep) return fail(PB_ERR_INVALID, "NULL argument");
    if (e->pending) {
        CK(cudaSetDevice(e->device));
        CK(cudaStreamSynchronize(e->stream));
        const ReportDev& r = *e->h_rep;
        pb_report o{};
        o.particle_steps = (int64_t)r.particle_steps;
        o.n_error = (int64_t)r.n_error;
        o.n_deleted = (int64_t)r.n_deleted;
        o.first_error_iter = r.first_error_iter == LLONG_MAX ? -1 : r.first_error_iter;
        o.n_out_of_time = (int64_t)r.n_out_of_time;
        o.max_iters_done = r.max_iters_done;
        o.cache_refills = (int64_t)r.cache_refills;
        o.max_state = r.max_state;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        o.kernel_ms = ms;
        e->last = o;
        e->pending = false;
    }
    *rep = e->last;
    return PB_OK;
}

int32_t pb_advect(pb_engine* e, const pb_advect_args* a, pb_report* rep) {
    int32_t rc = pb_advect_async(e, a);
    if (rc) return rc;
    pb_report tmp;
    rc = pb_last_report(e, &tmp);
    if (rc) return rc;
    if (rep) *rep = tmp;
    return PB_OK;
}

int32_t pb_flag_view_outside_time(pb_engine* e, double dt, double endtime) {
    if (!e) return fail(PB_ERR_INVALID, "engine is NULL");
    CK(cudaSetDevice(e->device));
    if (e->n > 0) {
        ParticlesDev P{(float*)e->px.p, (float*)e->py.p, (float*)e->pz.p, (float*)e->pdx.p, (float*)e->pdy.p,
                       (float*)e->pdz.p, (double*)e->pt.p, (int*)e->pstate.p, (int*)e->pei.p, (long long*)e->ppid.p, e->n};
        flag_view_kernel<<<(unsigned)((e->n + 255) / 256), 256, 0, e->stream>>>(P, dt, endtime);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(e->stream));
    return PB_OK;
}

int32_t pb_debug_normals(pb_engine* e, uint64_t seed, uint64_t rng_call, int64_t iter, int64_t n,
                         const int64_t* particle_id, double* out) {
    if (!e || (n && (!particle_id || !out))) return fail(PB_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(e->device));
    if (n == 0) return PB_OK;
    long long* d_pid = nullptr;
    double* d_out = nullptr;
    CK(cudaMalloc(&d_pid, n * 8));
    CK(cudaMalloc(&d_out, n * 16));
    CK(cudaMemcpyAsync(d_pid, particle_id, n * 8, cudaMemcpyHostToDevice, e->stream));
    normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(seed, rng_call, iter, n, d_pid, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, n * 16, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_pid);
    cudaFree(d_out);
    return PB_OK;
}

}  // extern "C"
