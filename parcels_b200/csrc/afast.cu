// afast.cu -- the headline hot path: AdvectionRK4 / AdvectionRK4_3D with XLinear_Velocity (reference kernels/_advection.py:42-75,
// interpolators/_xinterpolators.py:78-190, _core/field.py:250-405) on a rectilinear A-grid with FLOAT64 coordinates, FLOAT32
// data and a time axis, every level resident (BASELINE configs 2, 4, 5 and the north-star workload).  Same arithmetic, operation
// by operation, as the generic AGridPolicy<double, float, true, NC, 0> of agrid.cuh (which stays the kernel of every other
// dtype / scheme / interpolator combination and is the cross-check of this one in the parity tests); what differs is the schedule:
//
//  * ONE evaluation site in a 4-trip stage loop.  The common case -- the sample is still inside the cached cell of every axis --
//    is straight-line code: three bcoord divisions, the cos of the latitude, Z-lerp, bilinear, unit conversion.  Everything else
//    (a cell change, a sentinel index, a sample on the first node of an axis, outside the time axis) is a rare side path.
//  * The T-lerped block (8 values per component, float64) is kept per lane next to the raw 2x2x2x2 block (float32): the
//    sample time of stage 3 equals stage 2's and stage 1's equals the previous step's stage 4, so the 24 time lerps are
//    done twice per step instead of four times (odd stages renew, even stages reuse -- a warp-uniform decision).
//  * Per-lane cache in shared memory as 16-byte columns ([chunk][lane]: LDS.128 / STS.128, conflict-free): 12 x float4 raw +
//    12 x double2 lerped = 384 B per lane, 48 KB per 128-thread block, 4 blocks per SM.
//  * A refill reads the NODE-INTERLEAVED copy of U, V, W ({u, v, w, 0} per node, built on the device at upload,
//    interleave_kernel below): 16 x 16-byte loads instead of 48 scattered 4-byte ones, a third of the DRAM sectors.
//  * On a cell miss the two neighbouring cells are tried before the full binary search (a particle crosses one face at a time).
#include "afast.cuh"

// ------------------------------------------------------------------------------------------------
// node-interleaved field copy
// ------------------------------------------------------------------------------------------------
__global__ void interleave_kernel(const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w, long long n,
                                  float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 r;
    r.x = u[i];
    r.y = v[i];
    r.z = w ? w[i] : 0.f;
    r.w = 0.f;
    out[i] = r;
}

cudaError_t launch_interleave(const float* u, const float* v, const float* w, long long nodes, void* out, cudaStream_t s) {
    if (nodes <= 0) return cudaSuccess;
    interleave_kernel<<<(unsigned)((nodes + 255) / 256), 256, 0, s>>>(u, v, w, nodes, (float4*)out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// per-lane state
// ------------------------------------------------------------------------------------------------
// Everything a lane caches lives in SHARED MEMORY as 16-byte columns [chunk][lane] (LDS.128 / STS.128, conflict-free):
//   raw block (float32)            NC * NV / 4 chunks
//   T-lerped block (float64)       NC * NL / 2 chunks
//   cells: {lo, hi} of z, y, x, t  4 chunks      -- read by the hit test of every evaluation, written by the side path only
//   reciprocals of the cell widths 2 chunks      -- {1 / dz, 1 / dy}, {1 / dx, 1 / dt}, for div_by_cached
// = 480 B per lane for UVW.  Keeping the cells out of the registers matters twice: the side path (which rewrites them) no longer
// forces register-to-register copies onto the hit path (measured: 80 moves per evaluation, profiles/README.md r02d), and the
// kernel's loop-carried state fits the register file without spills.  One 480-thread block per SM (15 warps).
// A cell that the straight-line path may use has lo < hi finite and idx >= 0; after a search that ended on a sentinel
// (or on the first node of the axis, bcoord == 0) lo = hi = NaN so that `x > lo && x <= hi` fails and the lane takes the
// side path again at its next evaluation.
struct FastCtx {
    int ti, zi, yi, xi;       // cell indices of the last search (-100: none): key of the raw block, and what `ei` is raveled from
    double lerp_t;            // sample time the T-lerped block in shared memory belongs to (-1: none)
    float4* raw;              // this lane's column of 16-byte chunks (raw block first, see the layout above)
    bool searched;
    int state;
    int ei;
    unsigned int refills;
    bool out_of_time;
    signed char len_t, len_z;  // (interface of the generic kernel skeleton; unused here)
};

// NC: components sampled (2: fieldset.UV, 3: fieldset.UVW).  HZ: the grid has a depth axis with >= 2 levels (else zi = 0,
// zeta = 0 and there is no Z-lerp: `lenZ == 1`, _xinterpolators.py:131).
template <int NC_, bool HZ>
struct AFastPolicy {
    static constexpr int NC = NC_;
    static constexpr int NV = HZ ? 16 : 8;        // raw values per component: (t, [z,] y, x) corners
    static constexpr int NL = NV / 2;             // T-lerped values per component
    static constexpr int RAW_CHUNKS = NC * NV / 4;  // float4 columns
    static constexpr int LRP_CHUNKS = NC * NL / 2;  // double2 columns
    static constexpr int CELL_CHUNKS = 4;           // double2 columns {lo, hi}: z, y, x, t
    static constexpr int RCP_CHUNKS = 2;            // double2 columns: {1 / dz, 1 / dy}, {1 / dx, 1 / dt} of the current cells
    static constexpr size_t SMEM = (size_t)(RAW_CHUNKS + LRP_CHUNKS + CELL_CHUNKS + RCP_CHUNKS) * 16 * PB_FAST_BLOCK;
    static constexpr bool RUNTIME_DTYPE = false;
    static constexpr bool FAST_RK4 = true;
    static constexpr bool FAST_UNROLLED = false;
    static constexpr bool F32_STAGES = false;
    static constexpr bool BATCH_LEN_T = false;  // float64 grid: every barycentric coordinate is float64, lenT changes no dtype
    static constexpr bool BATCH_LEN_Z = false;
    using Ctx = FastCtx;

    __device__ static __forceinline__ double2* lrp(const Ctx& e) {
        return reinterpret_cast<double2*>(e.raw + (size_t)RAW_CHUNKS * PB_FAST_BLOCK);
    }
    // cell k (0: z, 1: y, 2: x, 3: t) as {lo, hi}
    __device__ static __forceinline__ double2* cell(const Ctx& e, int k) {
        return reinterpret_cast<double2*>(e.raw + (size_t)(RAW_CHUNKS + LRP_CHUNKS + k) * PB_FAST_BLOCK);
    }
    // reciprocals of the cell widths (correctly rounded: computed by a division when the cell changes), for div_by_cached
    __device__ static __forceinline__ double2* rcp(const Ctx& e) {
        return reinterpret_cast<double2*>(e.raw + (size_t)(RAW_CHUNKS + LRP_CHUNKS + CELL_CHUNKS) * PB_FAST_BLOCK);
    }

    __device__ static __forceinline__ void init(Ctx& e, const AdvectParams&, int ei) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        e.ti = e.zi = e.yi = e.xi = -100;
        extern __shared__ __align__(16) unsigned char pb_smem[];
        e.raw = reinterpret_cast<float4*>(pb_smem) + threadIdx.x;
        double2 d;
        d.x = nan; d.y = nan;
#pragma unroll
        for (int k = 0; k < CELL_CHUNKS; ++k) *cell(e, k) = d;
        e.lerp_t = -1.0;  // valid sample times are >= 0
        e.ei = ei;
        e.searched = false;
        e.len_t = e.len_z = -1;
    }
    // ravel_index (basegrid.py:259-278) over the axes present of the last completed search; int64 arithmetic stored to int32
    __device__ static __forceinline__ void finish(Ctx& e, const AdvectParams& p) {
        if (!e.searched) return;
        int gxi = e.xi;
        if (p.g.decomposed && gxi >= 0) gxi += p.g.xi_offset;  // mode D: local column -> global column
        long long r = (long long)e.yi * p.g.xdim + (long long)gxi;
        if (p.g.nz > 0) r += (long long)(HZ ? e.zi : 0) * (p.g.ydim * p.g.xdim);
        e.ei = (int)r;
    }

    // gather the (t, z, y, x) corner block of every component from the node-interleaved copy into this lane's raw columns:
    // per (t, z) plane four 16-byte node loads -> one float4 chunk per component
    __device__ static __forceinline__ void refill(const FieldDev& f, Ctx& e, int ti, int zi, int yi, int xi) {
        const long long ot[2] = {wrap_idx(ti, f.T) * f.sT, up_idx(ti, f.T) * f.sT};
        const long long oz[2] = {wrap_idx(zi, f.Z) * f.sZ, up_idx(zi, f.Z) * f.sZ};
        const long long oy[2] = {wrap_idx(yi, f.Y) * f.sY, up_idx(yi, f.Y) * f.sY};
        const long long ox[2] = {wrap_idx(xi, f.X) * f.sX, up_idx(xi, f.X) * f.sX};
        const float4* __restrict__ base = (const float4*)f.il;
#pragma unroll
        for (int pl = 0; pl < NV / 4; ++pl) {
            // chunk pl of a component holds k = 4 pl .. 4 pl + 3,  k = (t * 2 + z) * 4 + y * 2 + x (HZ) | t * 4 + y * 2 + x
            const long long o = ot[HZ ? (pl >> 1) : pl] + (HZ ? oz[pl & 1] : 0);
            const float4 n00 = ldg(base + o + oy[0] + ox[0]), n01 = ldg(base + o + oy[0] + ox[1]);
            const float4 n10 = ldg(base + o + oy[1] + ox[0]), n11 = ldg(base + o + oy[1] + ox[1]);
            float4 q;
            q.x = n00.x; q.y = n01.x; q.z = n10.x; q.w = n11.x;
            e.raw[(0 * (NV / 4) + pl) * PB_FAST_BLOCK] = q;
            q.x = n00.y; q.y = n01.y; q.z = n10.y; q.w = n11.y;
            e.raw[(1 * (NV / 4) + pl) * PB_FAST_BLOCK] = q;
            if (NC == 3) {
                q.x = n00.z; q.y = n01.z; q.z = n10.z; q.w = n11.z;
                e.raw[(2 * (NV / 4) + pl) * PB_FAST_BLOCK] = q;
            }
        }
        e.refills++;
    }

    // Z-lerp (:141-145) and bilinear (:147-152, left to right) of one component's T-lerped values
    __device__ static __forceinline__ double zxy(const double (&L)[NL], double zeta, double omz, double w00, double w01, double w10, double w11) {
        double r[4];
        if (HZ) {
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = L[j] * omz + L[4 + j] * zeta;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = L[j];
        }
        return w00 * r[0] + w01 * r[1] + w10 * r[2] + w11 * r[3];
    }

    // One VectorField.eval (field.py:250-304) at an RK4 stage position; k = stage 0..3 (warp-uniform).
    __device__ static __forceinline__ void eval_fast(const AdvectParams& p, Ctx& e, const int k, const double ts, const double zs,
                                                     const double ys, const double xs, double& u, double& v, double& w) {
        const GridDev& g = p.g;
        const FieldDev& f = p.f;
        const bool renew = (k & 1) != 0;  // odd stages sample a new time: the T-lerped block is renewed; even stages reuse it
        // The cells are READ FROM SHARED MEMORY at every evaluation, on purpose: without this barrier the compiler keeps them in
        // registers across evaluations (they only change in the side path) and pays for it with ~80 register moves per evaluation
        // on the hit path plus spills of the particle state (volatile loads: neither NVVM nor ptxas may forward them).
        bool lerp_now = renew, retried = false;
        double2 bz, by, bx, bt;  // {lo, hi} of the current cells
        // At most two trips: the hit test; on a miss the side path (which rewrites the cells IN SHARED MEMORY) and the test again.
        // The cell values used below have ONE definition -- these loads -- so no register copies are needed where the paths join.
        for (;;) {
            bz = HZ ? lds_volatile(cell(e, 0)) : double2{0.0, 0.0};
            by = lds_volatile(cell(e, 1));
            bx = lds_volatile(cell(e, 2));
            bt = (renew || retried) ? lds_volatile(cell(e, 3)) : double2{0.0, 0.0};  // (even stages test the cached lerp's time)
            bool hit = xs > bx.x && xs <= bx.y && ys > by.x && ys <= by.y;
            if (HZ) hit = hit && zs > bz.x && zs <= bz.y;
            hit = hit && (renew ? (ts > bt.x && ts <= bt.y) : (ts == e.lerp_t));
            if (hit || retried) break;
            // ---------------- side path: searches, states, refill; special samples are finished here ----------------
            if (!(0 <= ts && ts <= g.time_len)) {  // OutsideTimeInterval (index_search.py:85-86): state 70, sample (0, 0, 0)
                e.state = PB_ERROR_OUTSIDE_TIME_INTERVAL;
                e.out_of_time = true;
                u = v = w = 0.0;
                return;
            }
            if (!renew) bt = lds_volatile(cell(e, 3));
            AxisCell<double> ct{e.ti, bt.x, bt.y}, cz{e.zi, bz.x, bz.y}, cy{e.yi, by.x, by.y}, cx{e.xi, bx.x, bx.y};
            const int oti = ct.idx, ozi = cz.idx, oyi = cy.idx, oxi = cx.idx;  // key of the raw block
            const double wt = ct.hi - ct.lo, wz = cz.hi - cz.lo, wy = cy.hi - cy.lo, wx = cx.hi - cx.lo;  // (NaN for a poisoned cell)
            if (oti >= 0 && !(ct.lo == ct.lo)) ct.idx = -100;  // (a poisoned cell is not a neighbour-search seed)
            if (ozi >= 0 && !(cz.lo == cz.lo)) cz.idx = -100;
            if (oyi >= 0 && !(cy.lo == cy.lo)) cy.idx = -100;
            if (oxi >= 0 && !(cx.lo == cx.lo)) cx.idx = -100;
            axis_locate(g.time, g.nt, ts, ct);
            const int ti = ct.idx;
            int zi = 0;
            if (HZ) {
                axis_locate((const double*)g.depth, g.nz, zs, cz);
                zi = cz.idx;
            }
            axis_locate((const double*)g.lat, g.ny, ys, cy);
            axis_locate((const double*)g.lon, g.nx, xs, cx);
            const int yi = cy.idx, xi = cx.idx;
            {   // cell widths changed: new reciprocals (the division itself: correctly rounded)
                double2* const rp = rcp(e);
                const double nwt = ct.hi - ct.lo, nwz = cz.hi - cz.lo, nwy = cy.hi - cy.lo, nwx = cx.hi - cx.lo;
                // (one division per axis whose cell width changed -- a particle crosses one face at a time: usually one of the four)
                double* const r0 = reinterpret_cast<double*>(rp);
                double* const r1 = reinterpret_cast<double*>(rp + PB_FAST_BLOCK);
                if (HZ && !(nwz == wz)) r0[0] = 1.0 / nwz;
                if (!(nwy == wy)) r0[1] = 1.0 / nwy;
                if (!(nwx == wx)) r1[0] = 1.0 / nwx;
                if (!(nwt == wt)) r1[1] = 1.0 / nwt;
            }
            e.ti = ti; e.zi = zi; e.yi = yi; e.xi = xi;
            if (g.decomposed) {  // mode D: a sentinel at a slab edge that is not the edge of the global domain = halo too small
                if ((xi == -2 && !g.left_global) || (xi == -1 && !g.right_global)) e.state = max(e.state, 99);
            }
            e.searched = true;
            int s = e.state;
            if (xi == -1 || yi == -1 || zi == -1) s = max(s, (int)PB_ERROR_OUT_OF_BOUNDS);  // field.py:327-356
            if (zi == -2) s = max(s, (int)PB_ERROR_THROUGH_SURFACE);
            if (oti != ti || (HZ && ozi != zi) || oyi != yi || oxi != xi) {
                refill(f, e, ti, zi, yi, xi);
                e.lerp_t = -1.0;
            }
            // special: a sentinel index, or a sample not strictly inside (lo, hi] of the time / depth cell (tau or zeta == 0 on the
            // first node: lenT / lenZ == 1 for this particle, _xinterpolators.py:130-131; NaN; a degenerate cell)
            const bool t_in = ts > ct.lo && ts <= ct.hi, z_in = !HZ || (zs > cz.lo && zs <= cz.hi);
            const bool special = xi < 0 || yi < 0 || zi < 0 || !t_in || !z_in;
            {   // write the cells back; the ones a special sample involved are poisoned so that the lane comes back here next time
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                double2 d;
                d.x = ct.lo; d.y = ct.hi; if (!t_in) { d.x = nan; d.y = nan; }
                *cell(e, 3) = d;
                if (HZ) { d.x = cz.lo; d.y = cz.hi; if (zi < 0 || !z_in) { d.x = nan; d.y = nan; } *cell(e, 0) = d; }
                d.x = cy.lo; d.y = cy.hi; if (yi < 0) { d.x = nan; d.y = nan; }
                *cell(e, 1) = d;
                d.x = cx.lo; d.y = cx.hi; if (xi < 0) { d.x = nan; d.y = nan; }
                *cell(e, 2) = d;
            }
            if (special) {
                const double tau = axis_bcoord(g.nt, ts, ct), zeta = HZ ? axis_bcoord(g.nz, zs, cz) : 0.0;
                const double eta = axis_bcoord(g.ny, ys, cy), xsi = axis_bcoord(g.nx, xs, cx);
                const bool two_t = tau > 0;             // lenT, per particle (float64 grid: no dtype depends on the batch)
                const bool two_z = HZ && !(zeta <= 0);  // lenZ, per particle (a NaN depth must poison the value)
                u = special_component<NV>(e.raw, tau, zeta, eta, xsi, two_t, two_z);
                v = special_component<NV>(e.raw + (NV / 4) * PB_FAST_BLOCK, tau, zeta, eta, xsi, two_t, two_z);
                w = NC == 3 ? special_component<NV>(e.raw + 2 * (NV / 4) * PB_FAST_BLOCK, tau, zeta, eta, xsi, two_t, two_z) : 0.0;
                if (g.spherical) spherical(g, k, ys, u, v);
                if (u != u || v != v || w != w) s = max(s, (int)PB_ERROR_INTERPOLATION);
                if (xi < 0 || yi < 0 || zi < 0) { u = 0.0; v = 0.0; w = 0.0; }
                e.state = s;
                return;
            }
            e.state = s;
            lerp_now = true;  // (an even stage after a cell change: its block has to be lerped for this sample time first)
            retried = true;
        }
        // ---------------- straight-line path: every cell is current, 0 < bcoord <= 1 on every axis ----------------
        // bcoord = (x - lo) / (hi - lo), index_search.py:57 (denominator in the axis dtype), with the cell width's cached reciprocal
        const double2 r_zy = rcp(e)[0], r_xt = rcp(e)[PB_FAST_BLOCK];
        const double zeta = HZ ? div_by_cached(zs - bz.x, bz.y - bz.x, r_zy.x) : 0.0;
        const double eta = div_by_cached(ys - by.x, by.y - by.x, r_zy.y);
        const double xsi = div_by_cached(xs - bx.x, bx.y - bx.x, r_xt.x);
        const double omz = 1 - zeta;
        const double w00 = (1 - xsi) * (1 - eta), w01 = xsi * (1 - eta), w10 = (1 - xsi) * eta, w11 = xsi * eta;
        double2* const lp = lrp(e);
        double q[3] = {0.0, 0.0, 0.0};
        if (lerp_now) {
            const double tau = div_by_cached(ts - bt.x, bt.y - bt.x, r_xt.y);
            const double omt = 1 - tau;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float lo_t[NL], hi_t[NL];
#pragma unroll
                for (int j = 0; j < NL / 4; ++j) {
                    const float4 a = e.raw[(c * (NV / 4) + j) * PB_FAST_BLOCK];
                    const float4 b = e.raw[(c * (NV / 4) + NL / 4 + j) * PB_FAST_BLOCK];
                    lo_t[4 * j] = a.x; lo_t[4 * j + 1] = a.y; lo_t[4 * j + 2] = a.z; lo_t[4 * j + 3] = a.w;
                    hi_t[4 * j] = b.x; hi_t[4 * j + 1] = b.y; hi_t[4 * j + 2] = b.z; hi_t[4 * j + 3] = b.w;
                }
                double L[NL];
#pragma unroll
                for (int j = 0; j < NL; ++j) L[j] = (double)lo_t[j] * omt + (double)hi_t[j] * tau;  // _xinterpolators.py:135-139
#pragma unroll
                for (int j = 0; j < NL / 2; ++j) {
                    double2 d;
                    d.x = L[2 * j]; d.y = L[2 * j + 1];
                    lp[(c * (NL / 2) + j) * PB_FAST_BLOCK] = d;
                }
                q[c] = zxy(L, zeta, omz, w00, w01, w10, w11);
            }
            e.lerp_t = ts;
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                double L[NL];
#pragma unroll
                for (int j = 0; j < NL / 2; ++j) {
                    const double2 d = lp[(c * (NL / 2) + j) * PB_FAST_BLOCK];
                    L[2 * j] = d.x; L[2 * j + 1] = d.y;
                }
                q[c] = zxy(L, zeta, omz, w00, w01, w10, w11);
            }
        }
        u = q[0]; v = q[1]; w = NC == 3 ? q[2] : 0.0;
        if (g.spherical) spherical(g, k, ys, u, v);
        if (u != u || v != v || w != w) e.state = max(e.state, (int)PB_ERROR_INTERPOLATION);  // field.py:288-290
    }

    // u /= deg2m * cos(deg2rad(y)); v /= deg2m (_xinterpolators.py:182-184).  Stage 1 samples at the particle's own float32
    // latitude: the factor is float32 arithmetic there (and the float64 value is divided by its float64 promotion).
    __device__ static __forceinline__ void spherical(const GridDev& g, int k, double ys, double& u, double& v) {
        double conv;
        if (k == 0) conv = (double)((float)g.deg2m * cos_np(deg2rad_np((float)ys)));
        else conv = g.deg2m * cos_np(deg2rad_np(ys));
        u = u / conv;
        v = div_by_cached_guarded(v, g.deg2m, g.inv_deg2m);
    }

    // the generic skeleton's entry point is not used by a FAST_RK4 policy
    template <class PZ, class PY, class PX>
    __device__ static __forceinline__ void eval(const AdvectParams&, Ctx&, bool, double, PZ, PY, PX, Val&, Val&, Val&) {}
};

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
bool agrid_fast_applies(const AdvectParams& p, bool coord_f64, bool data_f64, bool has_time, int nc) {
    if (!coord_f64 || data_f64 || !has_time || !p.f.il || p.f.windowed || p.g.curvilinear) return false;
    if (!(p.scheme == PB_ADVECTION_RK4 || p.scheme == PB_ADVECTION_RK4_3D)) return false;
    if (nc == 3 && p.g.nz < 2) return false;
    return p.g.nt >= 2 && p.g.nx >= 2 && p.g.ny >= 2;
}

template <int NC, bool HZ, bool DIFF>
static cudaError_t launch_fast2(const AdvectParams& p, cudaStream_t s) {
    using Pol = AFastPolicy<NC, HZ>;
    const long long grid = (p.P.n + PB_FAST_BLOCK - 1) / PB_FAST_BLOCK;
    if (Pol::SMEM > 48 * 1024) {
        cudaError_t ce = cudaFuncSetAttribute(advect_kernel<Pol, DIFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pol::SMEM);
        if (ce != cudaSuccess) return ce;
    }
    advect_kernel<Pol, DIFF><<<(unsigned)grid, PB_FAST_BLOCK, Pol::SMEM, s>>>(p);
    return cudaGetLastError();
}
template <int NC, bool HZ>
static cudaError_t launch_fast1(const AdvectParams& p, cudaStream_t s) {
    return p.diffusion ? launch_fast2<NC, HZ, true>(p, s) : launch_fast2<NC, HZ, false>(p, s);
}

cudaError_t launch_agrid_fast(const AdvectParams& p, int nc, cudaStream_t s) {
    const bool hz = p.g.nz >= 2;
    if (nc == 3) return launch_fast1<3, true>(p, s);
    return hz ? launch_fast1<2, true>(p, s) : launch_fast1<2, false>(p, s);
}
