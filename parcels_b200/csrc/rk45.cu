// rk45.cu -- AdvectionRK45 (reference kernels/_advection.py:85-155) inside Kernel.execute's loop with the
// Repeat / next_dt state machine (reference _core/kernel.py:108-120,190-245), one lane per particle, on
// rectilinear A-grids with XLinear_Velocity (AGridPolicy MODE 0 of agrid.cuh).
//
// What the reference does per loop iteration, restated per particle:
//   dt <- clamp(dt, time_to_endtime)                                   kernel.py:199-203 (per-particle dt array)
//   RK45(particle); while state == Repeat: RK45(particle) with dt / 2  kernel.py:206-216
//   x += dx; y += dy; z += dz; t += dt; dt <- next_dt                   kernel.py:108-120 (RK45 mode: fieldset.RK45_tol)
//   dt is NOT reverted to the nominal step                              kernel.py:224-226
// RK45 itself ends with `state = where(good, Evaluate, state)` / `state = where(~good, Repeat, state)`
// (_advection.py:140,154): every error state raised by its six field evaluations is overwritten, so a pure
// RK45 list never errors or deletes -- out-of-bounds samples are simply zero velocity.
#ifndef PB_SMEM_CACHE
#define PB_SMEM_CACHE
#endif
#include "agrid.cuh"
#include "rk45.cuh"


// The reference clamps the dt of EVERY particle at the top of every loop iteration (kernel.py:199-203), also of the
// ones that are not evaluated any more, and in RK45 mode never restores it.  A particle that left the loop after
// `iters` iterations while the batch went on to `total_iters` therefore ends with dt clamped against its final
// time_to_endtime (0 once it has reached endtime).
__global__ void rk45_finalize(ParticlesDev P, double* dt, const int* __restrict__ iters, long long total_iters, double endtime, int sign) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n || iters[i] >= total_iters) return;
    const double tte = sign * (endtime - P.t[i]);
    dt[i] = (sign == 1) ? fmax(fmin(dt[i], tte), 0.0) : fmin(fmax(dt[i], -tte), 0.0);
}

template <class A, class D, bool HT>
static cudaError_t launch1(const Rk45Params& q, cudaStream_t s) {
    using Pol = AGridPolicy<A, D, HT, 2, 0>;
    const size_t smem = (size_t)2 * 16 * sizeof(typename decltype(EvalCtx<A, D, 2>::cor)::S) * PB_BLOCK;
    if (smem > 48 * 1024) {
        cudaError_t ce = cudaFuncSetAttribute(rk45_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
    }
    rk45_kernel<Pol><<<(unsigned)((q.base.P.n + PB_BLOCK - 1) / PB_BLOCK), PB_BLOCK, smem, s>>>(q);
    return cudaGetLastError();
}

cudaError_t launch_rk45(const AdvectParams& p, double* dt, double* next_dt, int* iters, int next_dt_f32, double tol, double min_dt,
                        double max_dt, bool coord_f64, bool data_f64, bool has_time, cudaStream_t s) {
    Rk45Params q{p, dt, next_dt, iters, next_dt_f32, tol, min_dt, max_dt};
    if (coord_f64) {
        if (data_f64) return has_time ? launch1<double, double, true>(q, s) : launch1<double, double, false>(q, s);
        return has_time ? launch1<double, float, true>(q, s) : launch1<double, float, false>(q, s);
    }
    if (data_f64) return has_time ? launch1<float, double, true>(q, s) : launch1<float, double, false>(q, s);
    return has_time ? launch1<float, float, true>(q, s) : launch1<float, float, false>(q, s);
}

cudaError_t launch_rk45_finalize(const ParticlesDev& P, double* dt, const int* iters, long long total_iters, double endtime, int sign,
                                 cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    rk45_finalize<<<(unsigned)((P.n + 255) / 256), 256, 0, s>>>(P, dt, iters, total_iters, endtime, sign);
    return cudaGetLastError();
}
