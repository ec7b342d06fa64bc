// agrid.cuh -- A-grid policy of advect_kernel on rectilinear grids: per-lane 2x2x2x2 corner cache +
// NumPy-exact interpolation (reference interpolators/_xinterpolators.py:78-190, _core/field.py:250-405,
// _core/xgrid.py:316-356).  MODE selects the vector interpolator:
//   0  XLinear_Velocity (:169-190)                      -- instantiated in agrid.cu (the tuned hot path)
//   1  XFreeslip / XPartialslip (_Spatialslip, :385-495) -- instantiated in aslip.cu
//   2  XNearest per component (:515-560)                -- instantiated in aslip.cu
// and, with NC == 1, the scalar interpolators of Field.eval (_core/field.py:144-191) -- aslip.cu:
//   3  XLinear (:112-153)      4  XNearest (:515-560)      5  CGrid_Tracer (:335-383)      6  XLinearInvdistLandTracer (:556-613)
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// corner cache + XLinear (reference interpolators/_xinterpolators.py:78-153)
// ------------------------------------------------------------------------------------------------

// PB_SMEM_CACHE: keep the corner block in shared memory ([value][thread] layout: conflict-free, one
// column per lane) instead of registers -- frees 16*NC registers per thread for occupancy.
#ifndef PB_BLOCK
#define PB_BLOCK 128
#endif
// SF64: store the cached corner values converted to float64 (only for float64 grids, whose interpolation is
// float64 throughout -- a float32 value converts exactly): no F2F per read, twice the shared memory.
template <class D, int NC, bool SF64 = false>
struct Corners {
    int ti, zi, yi, xi;  // key of the block held in v (INT_MIN: empty)
#ifdef PB_SMEM_CACHE
    using S = typename std::conditional<SF64, double, D>::type;
    S* sm;               // this lane's column of the block-shared cache
    __device__ __forceinline__ void put(int c, int k, D val) { sm[(c * 16 + k) * PB_BLOCK] = (S)val; }
    __device__ __forceinline__ void load(int c, S (&out)[16]) const {
#pragma unroll
        for (int k = 0; k < 16; ++k) out[k] = sm[(c * 16 + k) * PB_BLOCK];
    }
    __device__ __forceinline__ S get(int c, int k) const { return sm[(c * 16 + k) * PB_BLOCK]; }  // k may be dynamic
#else
    using S = D;
    D v[NC][16];         // [component][(t*2+z)*4 + y*2 + x]
    __device__ __forceinline__ void put(int c, int k, D val) { v[c][k] = val; }
    __device__ __forceinline__ void load(int c, D (&out)[16]) const {
#pragma unroll
        for (int k = 0; k < 16; ++k) out[k] = v[c][k];
    }
#endif

    __device__ __forceinline__ void fill(const FieldDev& f, int nti, int nzi, int nyi, int nxi) {
        ti = nti; zi = nzi; yi = nyi; xi = nxi;
        long long ot[2] = {tslot(f, wrap_idx(nti, f.T)) * f.sT, tslot(f, up_idx(nti, f.T)) * f.sT};
        long long oz[2] = {wrap_idx(nzi, f.Z) * f.sZ, up_idx(nzi, f.Z) * f.sZ};
        long long oy[2] = {wrap_idx(nyi, f.Y) * f.sY, up_idx(nyi, f.Y) * f.sY};
        long long ox[2] = {wrap_idx(nxi, f.X) * f.sX, up_idx(nxi, f.X) * f.sX};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const D* __restrict__ base = (const D*)f.p[c];
#pragma unroll
            for (int k = 0; k < 16; ++k)
                put(c, k, ldg(base + ot[k >> 3] + oz[(k >> 2) & 1] + oy[(k >> 1) & 1] + ox[k & 1]));
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
__device__ __forceinline__ Val zlerp_bilinear(const C (&c)[8], TZ zeta, TY eta, TX xsi, bool two_z) {
    if (two_z) {
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
__device__ __forceinline__ Val xlinear(const D (&v)[16], TT tau, TZ zeta, TY eta, TX xsi, bool two_t, bool two_z) {
    if constexpr (std::is_same<TZ, double>::value && std::is_same<TY, double>::value && std::is_same<TX, double>::value) {
        // Every barycentric coordinate is float64 (float64 grid, or an RK stage position): all arithmetic
        // after the gather is float64 whatever D is, and a float32 corner value converts exactly.  One code
        // path; a skipped lerp just copies (x*(1-0) + y*0 == x would differ only for non-finite y).
        double r[8];
        if (two_t) {
            const double omt = 1 - (double)tau;
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = (double)v[k] * omt + (double)v[8 + k] * (double)tau;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = (double)v[k];
        }
        if (two_z) {
            const double omz = 1 - zeta;
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = r[k] * omz + r[4 + k] * zeta;
        }
        const double q = (1 - xsi) * (1 - eta) * r[0] + xsi * (1 - eta) * r[1] + (1 - xsi) * eta * r[2] + xsi * eta * r[3];
        return Val{q, false};
    } else {
        if (two_t) {
            using R = prom_t<D, TT>;
            R r[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = v[k] * (1 - tau) + v[8 + k] * tau;
            return zlerp_bilinear<R, TZ, TY, TX>(r, zeta, eta, xsi, two_z);
        } else {
            D r[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = v[k];
            return zlerp_bilinear<D, TZ, TY, TX>(r, zeta, eta, xsi, two_z);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// _Spatialslip helpers (reference _xinterpolators.py:385-480)
// ------------------------------------------------------------------------------------------------
// np.isclose(val, 0.0) on an array of dtype D: |val| <= 1e-8 with the tolerance cast to D (NaN: false)
template <class D, class S>
__device__ __forceinline__ bool near_zero(S val) {
    if constexpr (std::is_same<D, float>::value) return fabsf((float)val) <= 1e-8f;
    else return fabs((double)val) <= 1e-8;
}
// bit k = z*4 + y*2 + x set when the corner at the LOWER time level is ~0
template <class D, class S>
__device__ __forceinline__ unsigned zero_mask(const S (&blk)[16]) {
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) m |= near_zero<D, S>(blk[k]) ? (1u << k) : 0u;
    return m;
}
// value * factor with NumPy promotion (value dtype tracked in Val::f32, factor dtype F)
template <class F>
__device__ __forceinline__ Val mul_factor(Val q, F f) {
    if (q.f32 && std::is_same<F, float>::value) return Val{(double)((float)q.v * (float)f), true};
    return Val{q.v * (double)f, false};
}
// f = ones_like(own); f[land0 & (c > 0)] *= (a + b c) / c;  f[land1 & (c < 1)] *= (1 - b c) / (1 - c)   (:424-434)
// `own`: dtype of the factor array; C: dtype of the barycentric coordinate the damping acts along
template <class OWN, class C>
__device__ __forceinline__ OWN slip_factor(OWN f, bool land0, bool land1, C c, float a, float b) {
    if (land0 && c > 0) f = (OWN)(f * ((C)a + (C)b * c) / c);
    if (land1 && c < 1) f = (OWN)(f * (1 - (C)b * c) / (1 - c));
    return f;
}

// ------------------------------------------------------------------------------------------------
// per-particle evaluation state
// ------------------------------------------------------------------------------------------------
template <class A, class D, int NC>
struct EvalCtx {
    AxisCell<A> cx, cy, cz;
    AxisCell<double> ct;
#if defined(PB_SMEM_CACHE) && defined(PB_SMEM_F32)
    Corners<D, NC, false> cor;  // tuning variant: cached corners kept in the data dtype (half the shared memory, F2F per read)
#elif defined(PB_SMEM_CACHE)
    Corners<D, NC, std::is_same<A, double>::value> cor;
#else
    Corners<D, NC> cor;
#endif
    double last_t, last_tau;  // stages 2 and 3 of a step (and stage 4 / next stage 1) sample the same time
    int szi, syi, sxi;        // indices of the last completed search: ei is raveled once, on exit
    bool searched;
    int state;
    int ei;
    unsigned int refills;
    bool out_of_time;
    // lenT / lenZ of the reference are decided per BATCH (`any(tau > 0)`, _xinterpolators.py:130-131): -1 = decide per particle
    // (what a lane can know by itself), 0 / 1 = the batch's answer, given by whoever knows the batch (sampling pre-pass, host table)
    signed char len_t, len_z;
};

// VectorField.eval for one particle (reference _core/field.py:250-304,307-405 with
// XLinear_Velocity, _xinterpolators.py:169-190).  PZ/PY/PX: dtype of the sampled position
// (float32 = the particle's own arrays, float64 = an RK stage position).
template <class A, class D, bool HAS_TIME, int NC, int MODE, class PZ, class PY, class PX>
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
        if (t == e.last_t) {
            tau = (TT)e.last_tau;
        } else {
            tau = (TT)axis_search<double, double>(g.time, g.nt, t, e.ct);
            e.last_t = t;
            e.last_tau = (double)tau;
        }
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

    // -- particles.ei[:, igrid] = ravel_index(zi, yi, xi) (field.py:307-317): only the LAST eval's value
    //    survives, so the indices are kept and raveled once on exit (AGridPolicy::finish)
    int gxi = xi;
    if (g.decomposed) {
        // mode D: local column -> global column.  A sentinel at a slab edge that is NOT the edge of the global
        // domain means a stage position left owned+halo columns: the halo is too small for this dt.
        if (xi >= 0) gxi = xi + g.xi_offset;
        else if ((xi == -2 && !g.left_global) || (xi == -1 && !g.right_global)) e.state = max(e.state, 99);
    }
    e.szi = zi; e.syi = yi; e.sxi = gxi;
    e.searched = true;

    // -- state from positions (field.py:327-356).  X/Y index -2 is NOT an error in the reference.
    int s = e.state;
    if (xi == -1 || yi == -1 || zi == -1) s = max(s, (int)PB_ERROR_OUT_OF_BOUNDS);
    if (zi == -2) s = max(s, (int)PB_ERROR_THROUGH_SURFACE);

    // -- corner block: gather from HBM only when the bracketing block changed
    if (e.cor.ti != ti || e.cor.zi != zi || e.cor.yi != yi || e.cor.xi != xi) {
        e.cor.fill(f, ti, zi, yi, xi);
        e.refills++;
    }

    using DV = typename decltype(e.cor)::S;  // float64 copies on float64 grids (exact), else the data dtype
    // lenT / lenZ: the batch's decision when known, else the particle's own (zeta > 0 or NaN: a NaN depth must poison the value
    // like a batch-level lerp does)
    const bool two_t = HAS_TIME && (e.len_t < 0 ? (tau > 0) : (e.len_t != 0));  // (a field without a time axis: tau is all zeros)
    const bool two_z = e.len_z < 0 ? !(zeta <= 0) : (e.len_z != 0);
    if constexpr (MODE == 2 || MODE == 4 || MODE == 5) {
#ifdef PB_SMEM_CACHE
        // One node of the cached block per component, linear in time, no unit conversion.
        //   XNearest (_xinterpolators.py:515-560): the node on the near side of each axis (bcoord <= 0.5: lower);
        //     as a vector interpolator this is the reference's XNearest_Velocity (tests/test_interpolation.py:279).
        //   CGrid_Tracer (:335-383): the tracer point of the cell, index + SGRID offset (clipped like the block's
        //     upper corner; an index with a negative sentinel is masked to 0 afterwards, field.py:189).
        const int k = MODE == 5 ? (g.off_z ? 4 : 0) + (g.off_y ? 2 : 0) + (g.off_x ? 1 : 0)
                                : ((zeta <= (TZ)0.5) ? 0 : 4) + ((eta <= (TY)0.5) ? 0 : 2) + ((xsi <= (TX)0.5) ? 0 : 1);
        Val q[3] = {Val{0.0, false}, Val{0.0, false}, Val{0.0, false}};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const DV c0 = e.cor.get(c, k);
            if (two_t) q[c] = Val{(double)c0 * (1 - (double)tau) + (double)e.cor.get(c, 8 + k) * (double)tau, false};
            else q[c] = Val{(double)c0, std::is_same<D, float>::value};
        }
        u = q[0]; v = q[1];
        w = NC == 3 ? q[2] : Val{0.0, u.f32};
#endif
    } else if constexpr (MODE == 3 || MODE == 6) {
        DV blk[16];
        e.cor.load(0, blk);
        u = xlinear<DV, TT, TZ, TY, TX>(blk, tau, zeta, eta, xsi, two_t, two_z);  // scalar XLinear: the value as it is
        if constexpr (MODE == 6) {
            // XLinearInvdistLandTracer (:556-613): corners ~ 0 are land.  All gathered corners land -> 0; some land -> the
            // inverse-squared-distance mean of the ocean corners, distance in (eta, xsi) only, EVERY gathered time / depth
            // level summed alike (t, z, y, x order); a sample exactly on an ocean node takes the sum of that node over
            // the gathered levels.  The levels gathered are lenT x lenZ = (tau > 0) x (zeta > 0), per particle (DESIGN.md,
            // waiver 1).  `eta_b - j_grid` is a float array minus an int64 array: float64 whatever the bcoord dtype.
            const int nT = two_t ? 2 : 1, nZ = (e.len_z < 0 ? (zeta > 0) : two_z) ? 2 : 1;
            int n_land = 0;
            double num = 0.0, den = 0.0;
            D node_val = 0;  // np.where(exact_mask, corner_data, 0.0) keeps the DATA dtype: a float32 field sums in float32
            bool on_node = false;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if ((k >> 3) >= nT || ((k >> 2) & 1) >= nZ) continue;
                const DV c = blk[k];
                const double dj = (double)eta - (double)((k >> 1) & 1), di = (double)xsi - (double)(k & 1);
                const double d2 = dj * dj + di * di;
                if (near_zero<D, DV>(c)) {
                    ++n_land;
                } else {
                    const double inv = 1.0 / d2;
                    num = num + (double)c * inv;
                    den = den + inv;
                    if (d2 == 0) { node_val = node_val + (D)c; on_node = true; }
                }
            }
            const int n_all = nT * nZ * 4;
            double r = u.v;
            if (n_land == n_all) r = 0.0;
            else if (n_land > 0) r = on_node ? (double)node_val : num / den;
            u.v = u.f32 ? (double)(float)r : r;  // assigned into the XLinear result array: keeps its dtype
        }
        v = Val{0.0, u.f32}; w = v;
    } else {
        DV blk[16];
        unsigned land = 0;
        e.cor.load(0, blk);
        u = xlinear<DV, TT, TZ, TY, TX>(blk, tau, zeta, eta, xsi, two_t, two_z);
        if constexpr (MODE == 1) land = zero_mask<D, DV>(blk);
        e.cor.load(1, blk);
        v = xlinear<DV, TT, TZ, TY, TX>(blk, tau, zeta, eta, xsi, two_t, two_z);
        if constexpr (MODE == 1) {
            // _Spatialslip (:385-480): damp the velocity component parallel to a cell edge whose nodes are all land
            // (U and V ~ 0 at the lower time level).  The reference looks at the second depth level when ANY particle
            // of the batch has zeta > 0; per particle that is zeta > 0 (identical unless a batch mixes particles
            // exactly on a depth level with others).  W's factors always use both levels (:460-470).
            land &= zero_mask<D, DV>(blk);
            const unsigned l2 = (e.len_z < 0 ? (zeta > 0) : two_z) ? (land & (land >> 4)) : land;  // bit y*2+x: land on every level looked at
            const unsigned lw = land & (land >> 4);
            const float a = g.slip_a, b = g.slip_b;
            const TX f_u = slip_factor<TX, TY>((TX)1, (l2 & 3u) == 3u, (l2 & 12u) == 12u, eta, a, b);
            const TY f_v = slip_factor<TY, TX>((TY)1, (l2 & 5u) == 5u, (l2 & 10u) == 10u, xsi, a, b);
            u = mul_factor<TX>(u, f_u);
            v = mul_factor<TY>(v, f_v);
            if (g.spherical) {  // u /= 1852 * 60 * cos(deg2rad(y)); v /= 1852 * 60   (in-place: keeps the dtype)
                PY conv = (PY)111120 * cos_np(deg2rad_np(y));
                if (u.f32 && std::is_same<PY, float>::value) {
                    u.v = (double)((float)u.v / (float)conv);
                } else {
                    double q = u.v / (double)conv;
                    u.v = u.f32 ? (double)(float)q : q;
                }
                v.v = v.f32 ? (double)((float)v.v / 111120.0f) : v.v / 111120.0;
            }
            if (NC == 3) {
                e.cor.load(NC - 1, blk);
                w = xlinear<DV, TT, TZ, TY, TX>(blk, tau, zeta, eta, xsi, two_t, two_z);
                TZ f_w = slip_factor<TZ, TY>((TZ)1, (lw & 3u) == 3u, (lw & 12u) == 12u, eta, a, b);
                f_w = slip_factor<TZ, TX>(f_w, (lw & 5u) == 5u, (lw & 10u) == 10u, xsi, a, b);
                w = mul_factor<TZ>(w, f_w);
            } else {
                w = Val{0.0, u.f32};
            }
        } else {
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
                e.cor.load(NC - 1, blk);
                w = xlinear<DV, TT, TZ, TY, TX>(blk, tau, zeta, eta, xsi, two_t, two_z);
            } else {
                w = Val{0.0, u.f32};
            }
        }
    }
    // -- NaN -> ErrorInterpolation, then out-of-bounds samples -> 0 (field.py:288-290,359-378)
    if (u.v != u.v || v.v != v.v || w.v != w.v) s = max(s, (int)PB_ERROR_INTERPOLATION);
    if (xi < 0 || yi < 0 || zi < 0) {
        u.v = 0.0; v.v = 0.0; w.v = 0.0;
    }
    e.state = s;
}

template <class A, class D, bool HAS_TIME, int NC_, int MODE = 0>
struct AGridPolicy {
    static constexpr int NC = NC_;
    static constexpr bool RUNTIME_DTYPE = false;  // interpolation arithmetic is typed on the position dtype
    static constexpr bool FAST_RK4 = false;
    static constexpr bool F32_STAGES = (MODE == 2);  // nearest node: a stage value can be float32 at a float64 position
    static constexpr bool BATCH_LEN_T = std::is_same<A, float>::value;  // float32 grid: a value's dtype depends on the batch's lenT
    static constexpr bool BATCH_LEN_Z = (MODE == 1);  // _Spatialslip: the land test looks at lenZ depth levels
    using Ctx = EvalCtx<A, D, NC_>;
    __device__ static __forceinline__ void init(Ctx& e, const AdvectParams&, int ei) {
        e.cx.idx = e.cy.idx = e.cz.idx = e.ct.idx = -100;
        e.cx.lo = e.cx.hi = e.cy.lo = e.cy.hi = e.cz.lo = e.cz.hi = (A)0;
        e.ct.lo = e.ct.hi = 0.0;
        e.cor.ti = e.cor.zi = e.cor.yi = e.cor.xi = INT_MIN;
#ifdef PB_SMEM_CACHE
        extern __shared__ __align__(16) unsigned char pb_smem[];
        e.cor.sm = reinterpret_cast<typename decltype(e.cor)::S*>(pb_smem) + threadIdx.x;
#endif
        e.ei = ei;
        e.last_t = -1.0;  // valid sample times are >= 0
        e.last_tau = 0.0;
        e.searched = false;
        e.szi = e.syi = e.sxi = 0;
        e.len_t = e.len_z = -1;
    }
    // ravel_index (basegrid.py:259-278) over the axes present; int64 arithmetic stored to int32
    __device__ static __forceinline__ void finish(Ctx& e, const AdvectParams& p) {
        if (!e.searched) return;
        long long r = (long long)e.syi * p.g.xdim + (long long)e.sxi;
        if (p.g.nz > 0) r += (long long)e.szi * (p.g.ydim * p.g.xdim);
        e.ei = (int)r;
    }
    template <class PZ, class PY, class PX>
    __device__ static __forceinline__ void eval(const AdvectParams& p, Ctx& e, bool /*no_hint*/, double t, PZ z, PY y, PX x,
                                                Val& u, Val& v, Val& w) {
        eval_uvw<A, D, HAS_TIME, NC_, MODE, PZ, PY, PX>(p.g, p.f, e, t, z, y, x, u, v, w);
    }
};

