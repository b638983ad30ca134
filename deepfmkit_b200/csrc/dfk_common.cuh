// Shared definitions for the DFMI readout kernels (sm_100a).
// The numerical cores (Bessel recurrence, LM solver, EKF step) are written
// __host__ __device__ so that tests can also compile them with g++ and compare
// them with the oracle on a machine without a GPU; the shipped library only
// ever runs them on the device.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define DFK_HD __host__ __device__ __forceinline__
#define DFK_D __device__ __forceinline__
#else
#define DFK_HD inline
#endif

namespace dfk {

constexpr double kPi = 3.141592653589793;     // np.pi
constexpr double kTwoPi = 6.283185307179586;  // 2 * np.pi

// Mirror of dfk_lm_opts without the ABI padding concerns (passed by value to kernels).
struct LmOpts {
    int max_steps;
    double conv_improve, conv_param, fitok_threshold;
    double grid_min, grid_max, grid_step;
    double bessel_thr, sincos_thr;
};

struct LmCounts {
    unsigned long long n_state, n_ssq, n_solve, n_grid, n_bessel_steps;
};

DFK_HD void sincos_hd(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    ::sincos(x, s, c);
#else
    *s = std::sin(x);
    *c = std::cos(x);
#endif
}

}  // namespace dfk

#if defined(__CUDACC__)
#include <cuda_runtime.h>

namespace dfk {

// ---- warp helpers ------------------------------------------------------------------------
DFK_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) ---------------------------
DFK_D uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

DFK_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
DFK_D void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
DFK_D void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
DFK_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
DFK_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Streaming data is read exactly once: ask L2 to evict it first.
DFK_D uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
DFK_D void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}

}  // namespace dfk
#endif  // __CUDACC__
