// afast.cuh -- helpers shared by the two specialised RK4 kernels (afast.cu: one evaluation site in a stage loop; afast2.cu: the four
// stages written out, the side path out of line)
#pragma once
#include <cmath>

#include "agrid.cuh"

#ifndef PB_FAST_BLOCK
#define PB_FAST_BLOCK PB_BLOCK
#endif

// ------------------------------------------------------------------------------------------------
// axis search of the side path: the cell only (no bcoord division), neighbours of the cached cell first.
// Same cell / sentinel as axis_search (common.cuh; reference _core/index_search.py:20-62).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void axis_locate(const double* __restrict__ arr, int n, double x, AxisCell<double>& c) {
    if (n < 2) {  // index_search.py:45-46
        c.idx = 0;
        return;
    }
    if (c.idx >= 0) {
        // lo < x <= hi of a cell inside the axis is exactly searchsorted(side="left") - 1 == that cell
        if (x > c.lo && x <= c.hi) return;
        if (x > c.hi && c.idx + 2 <= n - 1) {
            const double nh = ldg(arr + c.idx + 2);
            if (x <= nh) { c.lo = c.hi; c.hi = nh; c.idx += 1; return; }
        } else if (x <= c.lo && c.idx >= 1) {
            const double nl = ldg(arr + c.idx - 1);
            if (x > nl) { c.hi = c.lo; c.lo = nl; c.idx -= 1; return; }
        }
    }
    int l = 0, h = n;  // first i with arr[i] >= x   (side="left")
    while (l < h) {
        const int m = (l + h) >> 1;
        if (ldg(arr + m) < x) l = m + 1; else h = m;
    }
    if (x != x) l = n;  // NaN sorts last
    const int i = min(max(l - 1, 0), n - 2);
    c.lo = ldg(arr + i);
    c.hi = ldg(arr + i + 1);
    c.idx = i;
    if (x < ldg(arr)) c.idx = -2;          // LEFT_OUT_OF_BOUNDS
    if (x > ldg(arr + n - 1)) c.idx = -1;  // RIGHT_OUT_OF_BOUNDS
}
// bcoord of x in the located cell (also of a clipped one: sentinels keep the cell of the nearest edge)
__device__ __forceinline__ double axis_bcoord(int n, double x, const AxisCell<double>& c) {
    return n < 2 ? 0.0 : (x - c.lo) / (c.hi - c.lo);
}

// The general evaluation of ONE component from a lane's raw block (special samples only: a sentinel index, a sample on the first
// node of the time / depth axis -- lenT or lenZ == 1 for this particle): agrid.cuh's xlinear, the all-float64 path.  Out of line:
// it is rare, and keeping its 16-value block out of the hot kernel's register allocation matters more than a call.
template <int NV>
__device__ __noinline__ double special_component(const float4* col, double tau, double zeta, double eta, double xsi, int two_t, int two_z) {
    float blk[16];
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
        const float4 r = col[j * PB_FAST_BLOCK];
        blk[4 * j] = r.x; blk[4 * j + 1] = r.y; blk[4 * j + 2] = r.z; blk[4 * j + 3] = r.w;
    }
    if (NV == 8) {  // (t, y, x) -> the generic (t, z, y, x) order with the one depth level twice
#pragma unroll
        for (int j = 3; j >= 0; --j) { blk[8 + j] = blk[4 + j]; blk[12 + j] = blk[4 + j]; }
#pragma unroll
        for (int j = 3; j >= 0; --j) blk[4 + j] = blk[j];
    }
    return xlinear<float, double, double, double, double>(blk, tau, zeta, eta, xsi, two_t != 0, two_z != 0).v;
}

// a 16-byte shared-memory load the compiler must issue where it is written (no hoisting, no store-to-load forwarding)
__device__ __forceinline__ double2 lds_volatile(const double2* p) {
#ifdef PB_HOSTSIM
    return *p;
#else
    double2 r;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
    return r;
#endif
}

