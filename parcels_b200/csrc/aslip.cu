// aslip.cu -- launchers of the other A-grid vector interpolators: XFreeslip / XPartialslip (MODE 1) and XNearest
// per component (MODE 2).  Device code in agrid.cuh; always built with the shared-memory corner cache.
#ifndef PB_SMEM_CACHE
#define PB_SMEM_CACHE
#endif
#include "agrid.cuh"
#include "rk45.cuh"

template <class A, class D, bool HT, int NC, int MODE>
static cudaError_t launch1(const AdvectParams& p, cudaStream_t s) {
    const int block = PB_BLOCK;
    const long long grid = (p.P.n + block - 1) / block;
#ifdef PB_SMEM_CACHE
    const size_t smem = (size_t)NC * 16 * sizeof(typename decltype(EvalCtx<A, D, NC>::cor)::S) * PB_BLOCK;
    if (smem > 48 * 1024) {
        cudaError_t ce = cudaFuncSetAttribute(advect_kernel<AGridPolicy<A, D, HT, NC, MODE>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
    }
#else
    const size_t smem = 0;
#endif
    advect_kernel<AGridPolicy<A, D, HT, NC, MODE>><<<(unsigned)grid, block, smem, s>>>(p);
    return cudaGetLastError();
}

template <class A, class D, int MODE>
static cudaError_t launch_ad(const AdvectParams& p, bool ht, int nc, cudaStream_t s) {
    if (ht) return nc == 3 ? launch1<A, D, true, 3, MODE>(p, s) : launch1<A, D, true, 2, MODE>(p, s);
    return nc == 3 ? launch1<A, D, false, 3, MODE>(p, s) : launch1<A, D, false, 2, MODE>(p, s);
}

template <int MODE>
static cudaError_t launch_mode(const AdvectParams& p, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s) {
    if (coord_f64) return data_f64 ? launch_ad<double, double, MODE>(p, has_time, nc, s) : launch_ad<double, float, MODE>(p, has_time, nc, s);
    return data_f64 ? launch_ad<float, double, MODE>(p, has_time, nc, s) : launch_ad<float, float, MODE>(p, has_time, nc, s);
}
cudaError_t launch_agrid_alt(const AdvectParams& p, int mode, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s) {
    return mode == 1 ? launch_mode<1>(p, coord_f64, data_f64, has_time, nc, s) : launch_mode<2>(p, coord_f64, data_f64, has_time, nc, s);
}

// AdvectionRK45 with XFreeslip / XPartialslip (rk45.cuh): 2-D, MODE 1
template <class A, class D, bool HT>
static cudaError_t rk45_slip1(const Rk45Params& q, cudaStream_t s) {
    using Pol = AGridPolicy<A, D, HT, 2, 1>;
    const size_t smem = (size_t)2 * 16 * sizeof(typename decltype(EvalCtx<A, D, 2>::cor)::S) * PB_BLOCK;
    if (smem > 48 * 1024) {
        cudaError_t ce = cudaFuncSetAttribute(rk45_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
    }
    rk45_kernel<Pol><<<(unsigned)((q.base.P.n + PB_BLOCK - 1) / PB_BLOCK), PB_BLOCK, smem, s>>>(q);
    return cudaGetLastError();
}
cudaError_t launch_rk45_slip(const AdvectParams& p, double* dt, double* next_dt, int* iters, int next_dt_f32, double tol, double min_dt,
                             double max_dt, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s) {
    Rk45Params q{p, dt, next_dt, iters, next_dt_f32, tol, min_dt, max_dt};
    if (coord_f64) {
        if (data_f64) return has_time ? rk45_slip1<double, double, true>(q, s) : rk45_slip1<double, double, false>(q, s);
        return has_time ? rk45_slip1<double, float, true>(q, s) : rk45_slip1<double, float, false>(q, s);
    }
    if (data_f64) return has_time ? rk45_slip1<float, double, true>(q, s) : rk45_slip1<float, double, false>(q, s);
    return has_time ? rk45_slip1<float, float, true>(q, s) : rk45_slip1<float, float, false>(q, s);
}

template <class A, class D, bool HT, int NC, int MODE>
static cudaError_t sample1(const SampleParams& p, cudaStream_t s) {
#ifdef PB_SMEM_CACHE
    const size_t smem = (size_t)NC * 16 * sizeof(typename decltype(EvalCtx<A, D, NC>::cor)::S) * PB_BLOCK;
    if (smem > 48 * 1024) {
        cudaError_t ce = cudaFuncSetAttribute(sample_kernel<AGridPolicy<A, D, HT, NC, MODE>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
    }
#else
    const size_t smem = 0;
#endif
    sample_kernel<AGridPolicy<A, D, HT, NC, MODE>><<<(unsigned)((p.n + PB_BLOCK - 1) / PB_BLOCK), PB_BLOCK, smem, s>>>(p);
    return cudaGetLastError();
}
template <class A, class D, int MODE>
static cudaError_t sample_ad(const SampleParams& p, bool ht, int nc, cudaStream_t s) {
    if (ht) return nc == 3 ? sample1<A, D, true, 3, MODE>(p, s) : sample1<A, D, true, 2, MODE>(p, s);
    return nc == 3 ? sample1<A, D, false, 3, MODE>(p, s) : sample1<A, D, false, 2, MODE>(p, s);
}
template <int MODE>
static cudaError_t sample_mode(const SampleParams& p, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s) {
    if (coord_f64) return data_f64 ? sample_ad<double, double, MODE>(p, has_time, nc, s) : sample_ad<double, float, MODE>(p, has_time, nc, s);
    return data_f64 ? sample_ad<float, double, MODE>(p, has_time, nc, s) : sample_ad<float, float, MODE>(p, has_time, nc, s);
}
cudaError_t launch_sample_agrid_alt(const SampleParams& p, int mode, bool coord_f64, bool data_f64, bool has_time, int nc, cudaStream_t s) {
    return mode == 1 ? sample_mode<1>(p, coord_f64, data_f64, has_time, nc, s) : sample_mode<2>(p, coord_f64, data_f64, has_time, nc, s);
}

template <class A, class D, int MODE>
static cudaError_t scalar_ad(const SampleParams& p, bool ht, cudaStream_t s) {
    return ht ? sample1<A, D, true, 1, MODE>(p, s) : sample1<A, D, false, 1, MODE>(p, s);
}
template <int MODE>
static cudaError_t scalar_mode(const SampleParams& p, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s) {
    if (coord_f64) return data_f64 ? scalar_ad<double, double, MODE>(p, has_time, s) : scalar_ad<double, float, MODE>(p, has_time, s);
    return data_f64 ? scalar_ad<float, double, MODE>(p, has_time, s) : scalar_ad<float, float, MODE>(p, has_time, s);
}
cudaError_t launch_sample_scalar(const SampleParams& p, int mode, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s) {
    if (mode == 3) return scalar_mode<3>(p, coord_f64, data_f64, has_time, s);
    if (mode == 6) return scalar_mode<6>(p, coord_f64, data_f64, has_time, s);
    return mode == 4 ? scalar_mode<4>(p, coord_f64, data_f64, has_time, s) : scalar_mode<5>(p, coord_f64, data_f64, has_time, s);
}
