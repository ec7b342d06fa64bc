// common.cuh -- shared by the A-grid (agrid.cu) and C-grid (cgrid.cu) kernels and the host API
// (engine.cu): device descriptors, NumPy-compatible numeric helpers, the register-cached 1-D axis
// search, Philox, and the generic advect_kernel<Policy> (Kernel.execute's loop, one lane per particle).
#pragma once

#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <type_traits>

#include "../../include/parcels_b200.h"

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
    double inv_deg2m;  // RN(1 / deg2m), for div_by_cached
    double time_len;  // time[nt-1] - time[0]
    long long xdim, ydim, zdim;  // cell counts for ravel_index
    // curvilinear grids (lon/lat are 2-D (ny, nx) arrays) + the spatial hash of _core/spatialhash.py
    int curvilinear;
    int hash_bitwidth;
    const unsigned int* hkeys;    // unique Morton codes, ascending
    const long long* hstarts;     // CSR start of each key in hfaces
    const long long* hcounts;     // number of candidate faces per key
    const unsigned int* hfaces;   // flat face id j * (nx - 1) + i, ascending within a key
    long long hnkeys;
    const int* hbucket;           // hbucket[b] = first key index with (key >> hbucket_shift) >= b  (2^20 + 1 entries)
    int hbucket_shift;
    int pad3_;
    const unsigned long long* hqbox;  // per face: quantised bbox (the hash cells that list the face), 6 x 10 bits
    const double* cellproj;       // spherical curvilinear: per cell 16 doubles = pu[4], pv[4], eu[3], ev[3], pad[2]
    double hbox[6];               // xmin, xmax, ymin, ymax, zmin, zmax of the hash grid
    int off_x, off_y, off_z;      // C-grid staggering offsets (_xinterpolators.py:99-109)
    float slip_a, slip_b;         // _Spatialslip coefficients: XFreeslip (1, 0), XPartialslip (0.5, 0.5)
    // multi-GPU mode D (X-slab domain decomposition): this engine holds lon[xi_offset : xi_offset + nx]
    // of the global axis (owned columns + halo); cell indices written to `ei` are GLOBAL.
    int decomposed;
    int xi_offset;
    int left_global, right_global;  // whether the local lon edge is the edge of the global domain
    double own_lo, own_hi;          // particles with own_lo <= x < own_hi are advanced here, others migrate
};

struct FieldDev {
    const void* p[3];
    int T, Z, Y, X;            // data shape (shared by all components on an A-grid); T = levels of the time axis
    long long sT, sZ, sY, sX;  // element strides; 0 for size-1 (never indexed) dims
    // time-slab streaming (the GPU analogue of the reference's WindowedArray, _core/_windowed_array.py): the
    // buffers hold `ring` time levels, level L lives in slot L % ring; ring == T when every level is resident
    int ring;
    int windowed;
    double win_t0, win_t1;     // the resident levels cover sample times win_t0 <= t <= win_t1
    // node-interleaved copy of U, V, W (one 16-byte {u, v, w, 0} record per grid node, same (T, Z, Y, X) node order), built on the
    // device when float32 fields are uploaded: a corner-block refill of the hot kernel (afast.cu) is 16 vector loads instead of 48
    // scalar ones.  NULL when there is none (float64 data, time-windowed fields).
    const void* il;
};
__device__ __forceinline__ long long tslot(const FieldDev& f, long long level) { return f.windowed ? level % f.ring : level; }

struct ReportDev {
    unsigned long long particle_steps;
    unsigned long long n_error;
    unsigned long long n_deleted;
    long long first_error_iter;  // LLONG_MAX when none
    unsigned long long n_out_of_time;
    long long max_iters_done;
    unsigned long long cache_refills;
    unsigned long long n_migrate;  // mode D: particles that left the owned slab and wait for migration
    unsigned long long n_wait_window;  // time-slab streaming: particles whose next step needs a level not resident
    unsigned long long wait_t_min_bits, wait_t_max_bits;  // IEEE bits of min / max t of those particles (t >= 0)
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

// mode D, in-kernel migration over peer memory (NVLink): every rank owns a double-buffered INBOX in its HBM that its peers map
// (CUDA IPC); a lane whose particle left the owned slab claims a slot in the new owner's inbox with one system-scope atomic and
// stores the 48-byte record there itself -- the exchange runs inside the advection kernel, there is no pack pass and no
// all-to-all.  Layout of an inbox allocation: [header 0][header 1][cap records of slot 0][cap records of slot 1].
constexpr int PB_MIG_MAX_RANKS = 16;
constexpr int PB_MIG_HEADER_BYTES = 256;       // one header per slot: the arrival counter (+ padding to its own sectors)
constexpr int PB_STATE_MIGRATED = 97;          // engine-internal: the record now lives on another rank (dropped by pb_migrate_p2p_finish)
struct MigRecord {  // 48 B per migrating particle
    double t;
    long long pid;
    float x, y, z, dx, dy, dz;
    int state, ei;
};
struct MigDev {
    unsigned char* peer[PB_MIG_MAX_RANKS];  // base of every rank's inbox allocation as mapped into THIS process (own one included)
    const double* bounds;                   // nranks + 1 slab edges
    long long cap;                          // records per inbox slot
    int nranks, rank;
    int slot;                               // inbox slot of this round (0 / 1, flipped by every pb_migrate_p2p_finish)
    int on;
};
__host__ __device__ __forceinline__ unsigned long long* mig_counter(unsigned char* base, int slot) {
    return reinterpret_cast<unsigned long long*>(base + (size_t)slot * PB_MIG_HEADER_BYTES);
}
__host__ __device__ __forceinline__ MigRecord* mig_records(unsigned char* base, int slot, long long cap) {
    return reinterpret_cast<MigRecord*>(base + 2 * (size_t)PB_MIG_HEADER_BYTES + (size_t)slot * (size_t)cap * sizeof(MigRecord));
}

struct AdvectParams {
    GridDev g;
    FieldDev f;
    ParticlesDev P;
    MigDev mig;
    int scheme, diffusion, delete_on_error, kh_spherical;
    double dt, endtime, kh_zonal, kh_meridional, kh_deg2m;
    unsigned long long seed, rng_call;
    long long max_iters;
    int hint_all_zero;  // curvilinear: every hinted xi of the evaluated view is 0 => the reference skips
                        // the hint for the whole batch at the first eval (index_search.py:269-282)
    int resume;         // 1: continue a Kernel.execute call after a migration (states are NOT reset to Evaluate)
    int kernels_only;   // 1: one iteration's kernel functions only; the host does the position update etc.
    int batch_levels;  // PB_BATCH_* bits: what the reference decides per batch of an evaluation (lenT of the first one, lenZ)
    int lone_particle; // the whole set is ONE particle (not: this launch's chunk): it is its own batch for `if np.any(xi)` (cgrid.cu)
    int pad2_;
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
// np.cos of a float64 LATITUDE in radians (the unit conversion of _xinterpolators.py:182): cos(x) = sin(pi/2 - |x|) as ONE odd
// polynomial on [0, pi/2] -- no quadrant logic, no integer work: 16 float64 instructions where CUDA's cos() is ~60 of all
// kinds.  Max error 1.9 ulp against long double on 2.5e6 latitudes (CUDA documents 2 ulp for its own cos; NumPy's SIMD
// cos is <= 1 ulp): far inside the position tolerance of spherical meshes (2 float32 ulp).  Anything that is not a latitude
// (|x| > pi/2, NaN) takes cos().  PB_FAST_COS=0 builds with cos() for A/B runs.
static __device__ __noinline__ double cos_cold(double x) { return cos(x); }
// Taylor coefficients of sin(r) / r in r^2, highest first: -1/23!, 1/21!, ..., -1/3!.  In CONSTANT memory on purpose: as literals the
// compiler materialises each through two UMOVs (ncu, r02h: 13 % of the curvilinear kernel's executed instructions); from the
// constant bank they are an operand of the DFMA itself.
static __constant__ double PB_SIN_TAYLOR[11] = {-3.8681701706306841e-23, 1.9572941063391263e-20, -8.2206352466243295e-18, 2.8114572543455206e-15,
                                                -7.6471637318198164e-13, 1.6059043836821613e-10, -2.5052108385441720e-08, 2.7557319223985893e-06,
                                                -1.9841269841269841e-04, 8.3333333333333332e-03, -1.6666666666666666e-01};
// sin(r) for |r| <= pi/2: r + r^3 P(r^2)
__device__ __forceinline__ double sin_poly_pio2(double r) {
    const double s = r * r;
    double p = PB_SIN_TAYLOR[0];
#pragma unroll
    for (int k = 1; k < 11; ++k) p = fma(p, s, PB_SIN_TAYLOR[k]);
    return fma(r * s, p, r);
}
__device__ __forceinline__ double cos_np(double x) {
#if defined(PB_FAST_COS) && PB_FAST_COS == 0
    return cos(x);
#else
    const double ax = fabs(x);
    if (!(ax <= 1.5707963267948968)) return cos_cold(x);
    return sin_poly_pio2((1.5707963267948966 - ax) + 6.123233995736766e-17);  // pi/2 = hi + lo; hi - |x| is exact for |x| >= pi/4
#endif
}

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
    // counter = (pid_lo, pid_hi, iteration, call index lo); key = seed, with the HIGH word of the 64-bit call index folded into
    // key word 1 (mode D packs (call << 20) + round into it: distinct calls must never share a counter)
    philox4x32_10((uint32_t)pid, (uint32_t)((unsigned long long)pid >> 32), (uint32_t)iter, (uint32_t)rng_call,
                  (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(rng_call >> 32), r);
    // Box-Muller in FLOAT32 (round 2): the increment is a random number whose stream differs from NumPy's MT19937 anyway (DESIGN.md
    // waiver 3), so its last 29 bits buy nothing -- but the float64 log + sqrt + sincospi were ~300 of the ~750 instructions per
    // particle and step that DiffusionUniformKh adds (ncu, profiles/README.md r02f c4).  logf / sqrtf / sincospif are accurate to
    // 1-2 float32 ulp; u1 keeps its 2^-33 tail resolution (small counter values convert exactly: |z| reaches 6.8 as before).
    // Everything that USES the increment stays the reference's float64 arithmetic.
    const float two_m32 = 2.3283064365386963e-10f;  // 2^-32
    const float u1 = ((float)r[0] + 0.5f) * two_m32;
    const float u2 = ((float)r[1] + 0.5f) * two_m32;
    const float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    zx = (double)(rad * c);
    zy = (double)(rad * s);
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
// gather index helpers (NumPy fancy-index semantics of the reference's corner gathers)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long wrap_idx(int i, int n) {  // NumPy negative-index wrap
    int w = i < 0 ? i + n : i;
    return (long long)min(max(w, 0), n - 1);
}
__device__ __forceinline__ long long up_idx(int i, int n) {  // np.clip(i + 1, 0, n - 1)
    return (long long)min(max(i + 1, 0), n - 1);
}

// a / b, correctly rounded, for a divisor whose correctly rounded reciprocal r = RN(1 / b) is at hand (a cached cell width, a
// constant): q0 = RN(a r) is within 2 ulp of a / b, the first residual correction leaves it within 1/2 ulp + 2^-105 (faithful), and for
// a faithful q and r = RN(1 / b) Markstein's correcting step  q + (a - b q) r  (residual exact by FMA) rounds to RN(a / b) -- the
// same final step as the hardware-assisted division sequence, without its reciprocal refinement.  Explicit fma(): unaffected by
// -fmad=false.  Valid while no intermediate leaves the normal range; oracle/division_check.c checks 4e8 cases against `/`.
__device__ __forceinline__ double div_by_cached(double a, double b, double r) {
    double q = a * r;
    double e = fma(-b, q, a);
    q = fma(e, r, q);
    e = fma(-b, q, a);
    return fma(e, r, q);
}
// the same for numerators that come from DATA (velocities): zeros (land), non-finite values and values near the ends of the
// exponent range take the division itself
__device__ __forceinline__ double div_by_cached_guarded(double a, double b, double r) {
    const double m = fabs(a);
    if (m >= 1e-280 && m <= 1e280) return div_by_cached(a, b, r);
    return a == 0.0 ? a : a / b;
}

// u * 0.5 keeps u's dtype (Python float is weak)
__device__ __forceinline__ double half_of(const Val& a) { return a.f32 ? (double)((float)a.v * 0.5f) : a.v * 0.5; }

// ------------------------------------------------------------------------------------------------
// the kernel: Kernel.execute's loop, one lane per particle
// ------------------------------------------------------------------------------------------------
// Policy supplies: Ctx (with members state, ei, refills, out_of_time), init(Ctx&, params, ei), and
//   eval<PZ,PY,PX>(params, Ctx&, no_hint, t, z, y, x, u, v, w)  == VectorField.eval for one particle.
#ifndef PB_MINBLOCKS
#define PB_MINBLOCKS 3
#endif
#ifdef PB_BLOCK
#define PB_BLOCK_THREADS PB_BLOCK
#else
#define PB_BLOCK_THREADS 128
#endif
// DIFF = false compiles the kernel without the DiffusionUniformKh block (the specialised RK4 kernel is instantiated both ways:
// the advection-only hot path carries neither the Philox / Box-Muller code nor its loop-invariant registers)
// whether a policy is a FAST_RK4 one with the two-stage loop body (afast.cu, SCHED 2)
template <class P, class = void>
struct FastUnrolled : std::false_type {};
template <class P>
struct FastUnrolled<P, std::void_t<decltype(P::FAST_UNROLLED)>> : std::integral_constant<bool, P::FAST_UNROLLED> {};
template <class P>
constexpr bool fast_unrolled_v = FastUnrolled<P>::value;

template <class Policy, bool DIFF = true>
__global__ void __launch_bounds__(PB_BLOCK_THREADS, PB_MINBLOCKS) advect_kernel(const __grid_constant__ AdvectParams p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long my_steps = 0, my_refills = 0;  // (my_steps = the lane's iteration count, set on exit)
    int final_state = 0;
    long long my_iters = 0;
    bool errored = false, deleted = false, oot = false, migrate = false, wait_window = false;

    if (i < p.P.n) {
        float x = p.P.x[i], y = p.P.y[i], z = p.P.z[i];
        float dx = p.P.dx[i], dy = p.P.dy[i], dz = p.P.dz[i];
        double t = p.P.t[i];

        typename Policy::Ctx e;
        Policy::init(e, p, p.P.ei[i]);
        e.state = p.resume ? p.P.state[i] : (int)PB_EVALUATE;  // kernel.py:188
        e.refills = 0;
        e.out_of_time = false;
        if constexpr (Policy::BATCH_LEN_Z) {  // slip interpolators: the land test looks at the batch's depth levels
            if (p.batch_levels & PB_BATCH_TWO_Z) e.len_z = 1;
        }

        const int sign = p.dt > 0 ? 1 : -1;
        constexpr bool three_d = (Policy::NC == 3);  // RK4_3D / RK2_3D sample fieldset.UVW, the others fieldset.UV
        const int nstage = (p.scheme == PB_ADVECTION_NONE) ? 0 : (p.scheme == PB_ADVECTION_EE) ? 1 : ((p.scheme == PB_ADVECTION_RK2 || p.scheme == PB_ADVECTION_RK2_3D) ? 2 : 4);

        bool ei_zeroed = false;
        // DiffusionUniformKh: what does not change from step to step -- sqrt(|dt|) of the nominal step (the clamped last step takes
        // its own) and the meridional coefficient sqrt(2 Kh_meridional [/ deg2m^2]) -- is computed once (same operations, same values)
        [[maybe_unused]] double diff_sq = 0.0, diff_by = 0.0;
        if (DIFF && p.diffusion) {
            diff_sq = sqrt(fabs(p.dt));
            diff_by = sqrt(2 * (p.kh_spherical ? p.kh_meridional / (p.kh_deg2m * p.kh_deg2m) : p.kh_meridional));
        }
        long long it = 0;
        for (;; ++it) {
            if (p.max_iters >= 0 && it >= p.max_iters) break;
            const double tte = sign * (p.endtime - t);                       // kernel.py:191
            if (!((e.state == PB_SUCCESS || e.state == PB_EVALUATE) && tte >= 0)) break;  // :193-195
            if (p.g.decomposed && !((double)x >= p.g.own_lo && (double)x < p.g.own_hi)) {  // mode D: not ours (any more)
                migrate = true;
                break;
            }
            // adapt dt to end exactly on endtime (:199-203)
            const double dtp = (sign == 1) ? fmax(fmin(p.dt, tte), 0.0) : fmin(fmax(p.dt, -tte), 0.0);
            if (p.f.windowed) {  // every sample time of this step (t .. t+dt) must lie inside the resident time levels
                const double ta = t, tb = t + dtp;
                if (!(fmin(ta, tb) >= p.f.win_t0 && fmax(ta, tb) <= p.f.win_t1)) {
                    // outside the field's whole time interval is the reference's OutsideTimeInterval, not a wait
                    if (fmin(ta, tb) >= 0 && fmax(ta, tb) <= p.g.time_len) {
                        wait_window = true;
                        atomicMin(&p.rep->wait_t_min_bits, (unsigned long long)__double_as_longlong(t));  // rare: once per window slide
                        atomicMax(&p.rep->wait_t_max_bits, (unsigned long long)__double_as_longlong(t));
                        break;
                    }
                }
            }

            // ---- advection kernel (kernels/_advection.py) ----
            Val u1, v1, w1, uk, vk, wk;
            double su, sv, sw;  // running RK4 sums, left to right
            // F32_STAGES policies: the same sums in float32, valid while every term so far was float32
            [[maybe_unused]] bool xf32 = false, yf32 = false, zf32 = false;
            [[maybe_unused]] float sxf = 0.f, syf = 0.f, szf = 0.f;
            // with DiffusionUniformKh in the list, ei[:, -1] was overwritten with 0 for every particle by the
            // constant-field evals of the previous step: curvilinear hints are all zero again
            const bool nohint1 = (it == 0 && p.hint_all_zero) || (it > 0 && DIFF && p.diffusion);
            if constexpr (fast_unrolled_v<Policy>) {
                // afast.cu SCHED 2: a two-stage loop body (an even stage that reuses the T-lerped block, an odd one that renews it), run
                // twice per step; operation for operation the arithmetic of the four-trip loop below
                su = sv = sw = 0.0;
                uk = Val{0.0, false}; vk = uk; wk = uk;
                const double xd = (double)x, yd = (double)y, zd = (double)z;
                const double th = t + 0.5 * dtp;
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    const bool first = (j == 0);
                    // stage 1 (the particle's own position and time) or stage 3 (half a step with stage 2's velocities)
                    double xs = first ? xd : xd + (uk.v * 0.5) * dtp, ys = first ? yd : yd + (vk.v * 0.5) * dtp;
                    double zs = (first || !three_d) ? zd : zd + (wk.v * 0.5) * dtp;
                    Policy::template eval_fast<0>(p, e, false, first, first ? t : th, zs, ys, xs, uk.v, vk.v, wk.v);
                    if (first) { su = uk.v; sv = vk.v; sw = wk.v; }
                    else { su = su + 2.0 * uk.v; sv = sv + 2.0 * vk.v; sw = sw + 2.0 * wk.v; }
                    // stage 2 (half a step with stage 1's velocities) or stage 4 (a full step with stage 3's)
                    xs = xd + (first ? uk.v * 0.5 : uk.v) * dtp; ys = yd + (first ? vk.v * 0.5 : vk.v) * dtp;
                    zs = three_d ? zd + (first ? wk.v * 0.5 : wk.v) * dtp : zd;
                    Policy::template eval_fast<1>(p, e, true, false, first ? th : t + dtp, zs, ys, xs, uk.v, vk.v, wk.v);
                    const double m = first ? 2.0 : 1.0;  // u1 + 2*u2 + 2*u3 + u4
                    su = su + m * uk.v; sv = sv + m * vk.v; sw = sw + m * wk.v;
                }
            } else if constexpr (Policy::FAST_RK4) {
                // afast.cu SCHED 1: float64 grid, float32 data -- every stage value is float64 (Val::f32 never set), one eval site,
                // odd stages renew the T-lerped block, even stages reuse it (stages 2/3 and 4/next-1 share their sample time)
                su = sv = sw = 0.0;
                uk = Val{0.0, false}; vk = uk; wk = uk;
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const bool full = (k == 3);
                    const bool first = (k == 0);
                    const double xs = first ? (double)x : (double)x + (full ? uk.v : uk.v * 0.5) * dtp;
                    const double ys = first ? (double)y : (double)y + (full ? vk.v : vk.v * 0.5) * dtp;
                    const double zs = (first || !three_d) ? (double)z : (double)z + (full ? wk.v : wk.v * 0.5) * dtp;
                    const double ts = first ? t : t + (full ? dtp : 0.5 * dtp);
                    Policy::template eval_fast<-1>(p, e, (k & 1) != 0, first, ts, zs, ys, xs, uk.v, vk.v, wk.v);
                    if (first) { su = uk.v; sv = vk.v; sw = wk.v; }
                    else {
                        const double m = (k == 3) ? 1.0 : 2.0;  // u1 + 2*u2 + 2*u3 + u4
                        su = su + m * uk.v; sv = sv + m * vk.v; sw = sw + m * wk.v;
                    }
                }
            } else if constexpr (Policy::RUNTIME_DTYPE) {
                // One eval call site (the policy branches at run time where the position dtype matters):
                // keeps the heavy curvilinear search + C-grid code in the instruction cache.
                su = sv = sw = 0.0;
                uk = Val{0.0, false}; vk = uk; wk = uk;
#pragma unroll 1
                for (int k = 0; k < nstage; ++k) {
                    const bool full = (k == 3);
                    const bool first = (k == 0);
                    const double xs = first ? (double)x : (double)x + (full ? uk.v : half_of(uk)) * dtp;
                    const double ys = first ? (double)y : (double)y + (full ? vk.v : half_of(vk)) * dtp;
                    const double zs = (first || !three_d) ? (double)z : (double)z + (full ? wk.v : half_of(wk)) * dtp;
                    const double ts = first ? t : t + (full ? dtp : 0.5 * dtp);
                    if constexpr (Policy::BATCH_LEN_T) e.len_t = (first && it == 0 && (p.batch_levels & 1)) ? 1 : -1;
                    Policy::eval_rt(p, e, first && nohint1, ts, zs, ys, xs, /*xy_f32=*/first, /*z_f32=*/first || !three_d, uk, vk, wk);
                    if (first) { su = uk.v; sv = vk.v; sw = wk.v; }
                    else if (nstage == 4) {
                        const double m = (k == 3) ? 1.0 : 2.0;  // u1 + 2*u2 + 2*u3 + u4
                        su = su + m * uk.v; sv = sv + m * vk.v; sw = sw + m * wk.v;
                    }
                }
            } else {
            u1 = Val{0.0, false}; v1 = u1; w1 = u1;
            // float32 grids: the value dtype of this evaluation depends on the BATCH's lenT (only the call's first iteration can
            // hold a particle with tau == 0, i.e. exactly on the first time level; the host knows whether the others are too)
            if constexpr (Policy::BATCH_LEN_T) e.len_t = (it == 0 && (p.batch_levels & 1)) ? 1 : -1;
            if (nstage > 0) Policy::template eval<float, float, float>(p, e, nohint1, t, z, y, x, u1, v1, w1);
            if constexpr (Policy::BATCH_LEN_T) e.len_t = -1;
            su = u1.v; sv = v1.v; sw = w1.v;
            uk = u1; vk = v1; wk = w1;
            if constexpr (Policy::F32_STAGES) {
                xf32 = u1.f32; yf32 = v1.f32; zf32 = w1.f32;
                sxf = (float)u1.v; syf = (float)v1.v; szf = (float)w1.v;
            }
#ifdef PB_UNROLL_STAGES
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int k = 1; k < 4; ++k) {
                if (k >= nstage) break;
                // stage position: x + u*0.5*dt (k = 1, 2) or x + u*dt (k = 3)
                const bool full = (k == 3);
                const double xs = (double)x + (full ? uk.v : half_of(uk)) * dtp;
                const double ys = (double)y + (full ? vk.v : half_of(vk)) * dtp;
                const double ts = t + (full ? dtp : 0.5 * dtp);
                if constexpr (three_d) {
                    const double zs = (double)z + (full ? wk.v : half_of(wk)) * dtp;
                    Policy::template eval<double, double, double>(p, e, false, ts, zs, ys, xs, uk, vk, wk);
                } else {
                    Policy::template eval<float, double, double>(p, e, false, ts, z, ys, xs, uk, vk, wk);
                }
                if (nstage == 4) {
                    const double m = (k == 3) ? 1.0 : 2.0;  // u1 + 2*u2 + 2*u3 + u4
                    if constexpr (Policy::F32_STAGES) {
                        const float mf = (float)m;
                        if (xf32 && uk.f32) sxf = sxf + mf * (float)uk.v; else { if (xf32) su = (double)sxf; xf32 = false; }
                        if (yf32 && vk.f32) syf = syf + mf * (float)vk.v; else { if (yf32) sv = (double)syf; yf32 = false; }
                        if (zf32 && wk.f32) szf = szf + mf * (float)wk.v; else { if (zf32) sw = (double)szf; zf32 = false; }
                        // a float64 term: the float32 partial sum is promoted (exactly) and the addition is float64 from here on;
                        // a float32 term joining a float64 sum is promoted likewise (2 * float32 is exact in both)
                        if (!xf32) su = su + m * uk.v;
                        if (!yf32) sv = sv + m * vk.v;
                        if (!zf32) sw = sw + m * wk.v;
                    } else {
                        su = su + m * uk.v; sv = sv + m * vk.v; sw = sw + m * wk.v;
                    }
                }
            }
            }
            double ddx, ddy, ddz;
            if (nstage == 4) {
                if constexpr (Policy::F32_STAGES) {
                    // every stage value may be float32 (nearest-node sampling of float32 data at tau == 0 carries the DATA
                    // dtype whatever the position dtype): `(u1 + 2*u2 + 2*u3 + u4) / 6.0` is then float32 arithmetic
                    ddx = (xf32 ? (double)(sxf / 6.0f) : su / 6.0) * dtp;
                    ddy = (yf32 ? (double)(syf / 6.0f) : sv / 6.0) * dtp;
                    ddz = (zf32 ? (double)(szf / 6.0f) : sw / 6.0) * dtp;
                } else {
                    if constexpr (Policy::FAST_RK4) {  // (u1 + 2 u2 + 2 u3 + u4) / 6.0 with the constant's reciprocal: same quotient
                        constexpr double r6 = 1.0 / 6.0;
                        ddx = div_by_cached_guarded(su, 6.0, r6) * dtp;
                        ddy = div_by_cached_guarded(sv, 6.0, r6) * dtp;
                        ddz = three_d ? div_by_cached_guarded(sw, 6.0, r6) * dtp : 0.0;
                    } else {
                        ddx = su / 6.0 * dtp; ddy = sv / 6.0 * dtp; ddz = sw / 6.0 * dtp;
                    }
                }
            } else {  // EE: u1*dt ; RK2: u2*dt
                ddx = uk.v * dtp; ddy = vk.v * dtp; ddz = wk.v * dtp;
            }
            dx = (float)((double)dx + ddx);
            dy = (float)((double)dy + ddy);
            if (three_d) dz = (float)((double)dz + ddz);

            // ---- DiffusionUniformKh (kernels/_advectiondiffusion.py:120-153) ----
            if (DIFF && p.diffusion) {
                double zx, zy;
                wiener_normals(p.seed, p.rng_call, it, p.P.pid[i], zx, zy);  // (the id is read when needed: one register pair less in the loop)
                const double sq = dtp == p.dt ? diff_sq : sqrt(fabs(dtp));
                const double dWx = zx * sq, dWy = zy * sq;
                double khz = p.kh_zonal;
                if (p.kh_spherical) {
                    const float ang = (y * (float)3.14159265358979323846) / 180.0f;  // lat * np.pi / 180 in f32
                    const float m = (float)p.kh_deg2m * cosf(ang);
                    khz = khz / (double)(m * m);
                }
                const double bx = sqrt(2 * khz), by = diff_by;
                dx = (float)((double)dx + bx * dWx);
                dy = (float)((double)dy + by * dWy);
                ei_zeroed = true;  // the constant-field evals overwrite ei[:, -1] with cell 0 (model.py:292-318)
            }
            if (p.kernels_only) { ++it; break; }  // dx/dy/dz, state, ei are written back below; the host finishes the iteration
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
            if (e.state >= 50) { errored = true; ++it; break; }
        }
        Policy::finish(e, p);
        if (ei_zeroed) e.ei = 0;
        my_iters = it;
        my_steps = (unsigned long long)it;  // every counted iteration evaluated the kernels once (breaks before that do not count)
        my_refills = e.refills;
        oot = e.out_of_time;
        final_state = e.state;
        p.P.x[i] = x; p.P.y[i] = y; p.P.z[i] = z;
        p.P.dx[i] = dx; p.P.dy[i] = dy; p.P.dz[i] = dz;
        p.P.t[i] = t;
        p.P.state[i] = e.state;
        p.P.ei[i] = e.ei;
        if (migrate && p.mig.on) {
            // in-kernel migration: new owner = the slab that holds x (rank 0 owns (-inf, b1), the last rank [b_{n-1}, +inf));
            // a NaN x has no owner: it stays (and is not counted as a mover)
            const double xd = (double)x;
            int r = 0;
            while (r + 1 < p.mig.nranks && xd >= p.mig.bounds[r + 1]) ++r;
            if (xd != xd || r == p.mig.rank) migrate = false;
            else {
                unsigned char* const base = p.mig.peer[r];
                const unsigned long long slot = atomicAdd_system(mig_counter(base, p.mig.slot), 1ULL);
                if (slot < (unsigned long long)p.mig.cap) {  // (a full inbox: the particle waits here for the next round)
                    MigRecord rec;
                    rec.t = t; rec.pid = p.P.pid[i];
                    rec.x = x; rec.y = y; rec.z = z; rec.dx = dx; rec.dy = dy; rec.dz = dz;
                    rec.state = e.state; rec.ei = e.ei;
                    mig_records(base, p.mig.slot, p.mig.cap)[slot] = rec;
                    p.P.state[i] = PB_STATE_MIGRATED;
                }
            }
        }
    }

    // ---- report: warp-reduce then one atomic per warp ----
    const unsigned full = 0xffffffffu;
    unsigned long long s_steps = my_steps, s_ref = my_refills;
    unsigned n_err = errored, n_del = deleted, n_oot = oot, n_mig = migrate, n_ww = wait_window;
    // with the delete handler an out-of-interval sample is not an error of the lane, but the host still needs the iteration
    // it happened in (the reference deletes the WHOLE evaluated view of that iteration, field.py:31-44)
    long long mx_it = my_iters, mn_err = ((oot && deleted) || errored) ? my_iters - 1 : LLONG_MAX;  // the iteration the error arose in
    int mx_state = final_state;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_steps += __shfl_xor_sync(full, s_steps, o);
        s_ref += __shfl_xor_sync(full, s_ref, o);
        n_err += __shfl_xor_sync(full, n_err, o);
        n_del += __shfl_xor_sync(full, n_del, o);
        n_oot += __shfl_xor_sync(full, n_oot, o);
        n_mig += __shfl_xor_sync(full, n_mig, o);
        n_ww += __shfl_xor_sync(full, n_ww, o);
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
        if (n_mig) atomicAdd(&p.rep->n_migrate, (unsigned long long)n_mig);
        if (n_ww) atomicAdd(&p.rep->n_wait_window, (unsigned long long)n_ww);
        if (mx_it) atomicMax(&p.rep->max_iters_done, mx_it);
        if (mn_err != LLONG_MAX) atomicMin(&p.rep->first_error_iter, mn_err);
        if (mx_state) atomicMax(&p.rep->max_state, mx_state);
    }
}

// ------------------------------------------------------------------------------------------------
// VectorField.eval on arbitrary sample points (reference _core/field.py:250-304): one lane per sample.
// Used by the public FieldSet.UV/UVW.eval() of the host mirror and by the parity tests to pin a
// single evaluation (indices, state, velocities) against the oracle.
// ------------------------------------------------------------------------------------------------
struct SampleParams {
    GridDev g;
    FieldDev f;
    long long n;
    const double* t;
    const double* z;
    const double* y;
    const double* x;
    const int* ei_hint;  // may be NULL (no hint)
    int pos_f32;         // 1: positions are float32 values (the particle's own arrays), 0: float64
    int no_hint;
    double* u;
    double* v;
    double* w;
    int* ei_out;
    int* state_out;
    int* f32_out;        // may be NULL; 1 where NumPy's promotion gives the (first) value dtype float32
    // the reference's batch-level lenT / lenZ (`any(tau > 0)`, `any(zeta > 0)` over the samples of the call, _xinterpolators.py:130-131):
    // device word written by sample_flags_kernel before the sampling kernel runs: bit 0 = any(tau > 0), bit 1 = any(zeta > 0);
    // NULL = not known (every sample decides for itself)
    const int* batch_flags;
};

// The batch-level part of one Field.eval / VectorField.eval call: the time search and the depth search of every sample, exactly
// as the sampling kernel does them, reduced to `any(tau > 0)` and `any(zeta > 0)` (the reference's lenT / lenZ).
template <class A>
__global__ void sample_flags_kernel(const SampleParams s, int has_time, int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int m = 0;
    if (i < s.n) {
        const GridDev& g = s.g;
        if (has_time && g.nt > 0) {
            const double t = s.t[i];
            if (0 <= t && t <= g.time_len) {
                AxisCell<double> c;
                c.idx = -100; c.lo = c.hi = 0.0;
                if (axis_search<double, double>(g.time, g.nt, t, c) > 0) m |= 1;
            }
        }
        if (g.nz > 0) {
            AxisCell<A> c;
            c.idx = -100; c.lo = c.hi = (A)0;
            const double zeta = s.pos_f32 ? (double)axis_search<float, A>((const A*)g.depth, g.nz, (float)s.z[i], c)
                                          : (double)axis_search<double, A>((const A*)g.depth, g.nz, s.z[i], c);
            if (zeta > 0) m |= 2;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(flags, m);
}

template <class Policy>
__global__ void sample_kernel(const SampleParams s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    AdvectParams p{};
    p.g = s.g;
    p.f = s.f;
    typename Policy::Ctx e;
    Policy::init(e, p, s.ei_hint ? s.ei_hint[i] : 0);
    e.state = PB_EVALUATE;
    e.refills = 0;
    e.out_of_time = false;
    if (s.batch_flags) {  // the reference decides lenT / lenZ for the whole call
        const int bf = *s.batch_flags;
        e.len_t = (signed char)(bf & 1);
        e.len_z = (signed char)((bf >> 1) & 1);
    }
    Val u, v, w;
    const bool nh = s.no_hint || !s.ei_hint;
    if constexpr (Policy::RUNTIME_DTYPE) {
        const bool f = s.pos_f32 != 0;
        Policy::eval_rt(p, e, nh, s.t[i], f ? (double)(float)s.z[i] : s.z[i], f ? (double)(float)s.y[i] : s.y[i],
                        f ? (double)(float)s.x[i] : s.x[i], f, f, u, v, w);
    } else {
        if (s.pos_f32) Policy::template eval<float, float, float>(p, e, nh, s.t[i], (float)s.z[i], (float)s.y[i], (float)s.x[i], u, v, w);
        else Policy::template eval<double, double, double>(p, e, nh, s.t[i], s.z[i], s.y[i], s.x[i], u, v, w);
    }
    s.u[i] = u.v;
    if (s.v) s.v[i] = v.v;
    if (s.w) s.w[i] = w.v;
    if (s.f32_out) s.f32_out[i] = u.f32 ? 1 : 0;
    Policy::finish(e, p);
    s.ei_out[i] = e.ei;
    s.state_out[i] = e.state;
}

// launchers implemented in agrid.cu / cgrid.cu (one translation unit per grid family keeps nvcc parallel)
cudaError_t launch_agrid(const AdvectParams& p, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s);
// afast.cu: the specialised RK4 kernel (float64 grid, float32 node-interleaved data, time axis); `ok` tells whether it applies
bool agrid_fast_applies(const AdvectParams& p, bool coord_f64, bool data_f64, bool has_time, int nc);
cudaError_t launch_agrid_fast(const AdvectParams& p, int nc, int sched, cudaStream_t s);  // sched: 2 two-stage loop body, 1 four-trip loop
cudaError_t launch_interleave(const float* u, const float* v, const float* w, long long nodes, void* out, cudaStream_t s);
cudaError_t launch_cgrid(const AdvectParams& p, bool coord_f64, bool data_f64, int nc, cudaStream_t s);
// mode: 1 = _Spatialslip (g.slip_a/b), 2 = XNearest per component  (aslip.cu)
cudaError_t launch_agrid_alt(const AdvectParams& p, int mode, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s);
// AdvectionRK45 + the Repeat / next_dt state machine (rk45.cu); dt / next_dt / iters: per-particle device arrays
cudaError_t launch_rk45(const AdvectParams& p, double* dt, double* next_dt, int* iters, int next_dt_f32, double tol, double min_dt,
                        double max_dt, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s);
cudaError_t launch_rk45_slip(const AdvectParams& p, double* dt, double* next_dt, int* iters, int next_dt_f32, double tol, double min_dt,
                             double max_dt, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s);
cudaError_t launch_rk45_cgrid(const AdvectParams& p, double* dt, double* next_dt, int* iters, int next_dt_f32, double tol, double min_dt,
                              double max_dt, bool coord_f64, bool data_f64, cudaStream_t s);
cudaError_t launch_rk45_finalize(const ParticlesDev& P, double* dt, const int* iters, long long total_iters, double endtime, int sign,
                                 cudaStream_t s);
// device-side build of the curvilinear spatial-hash table from the per-face quantised boxes (hashbuild.cu)
struct HashTableDev {
    unsigned int* keys = nullptr;
    long long* starts = nullptr;
    long long* counts = nullptr;
    unsigned int* faces = nullptr;
    int* bucket = nullptr;
    long long nkeys = 0, nent = 0;
};
cudaError_t build_hash_table_device(const unsigned long long* d_qbox, long long nfaces, int bucket_bits, HashTableDev& t, cudaStream_t s);
// AdvectionDiffusionM1 (em = 0) / AdvectionDiffusionEM (em = 1) with the diffusivity fields fkz, fkm  (advdiff.cu)
cudaError_t launch_advdiff(const AdvectParams& p, const FieldDev& fkz, const FieldDev& fkm, int em, double dres, double deg2m_sq,
                           bool coord_f64, bool uv_f64, bool kh_f64, bool uv_time, bool kh_time, bool cgrid, cudaStream_t s);
// scalar Field.eval on a rectilinear grid, f.p[0] = the field: mode 3 XLinear, 4 XNearest, 5 CGrid_Tracer  (aslip.cu)
cudaError_t launch_sample_scalar(const SampleParams& p, int mode, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s);
// scalar Field.eval on a curvilinear grid: mode 4 XNearest, 5 CGrid_Tracer  (cgrid.cu)
cudaError_t launch_sample_scalar_curv(const SampleParams& p, int mode, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s);
cudaError_t launch_sample_agrid_alt(const SampleParams& p, int mode, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s);
cudaError_t launch_precompute_cells(const void* lon, const void* lat, int ny, int nx, bool coord_f64, double* out, cudaStream_t s);
cudaError_t launch_sample_agrid(const SampleParams& p, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s);
cudaError_t launch_sample_cgrid(const SampleParams& p, bool coord_f64, bool data_f64, int nc, cudaStream_t s);
// XLinear_Velocity behind the curvilinear search (curva.cu)
cudaError_t launch_curv_agrid(const AdvectParams& p, bool coord_f64, bool data_f64, int nc, cudaStream_t s);
cudaError_t launch_sample_curv_agrid(const SampleParams& p, bool coord_f64, bool data_f64, int nc, cudaStream_t s);
