// Shared device/host helpers for the zigma_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/zigma_b200.h"

// ---- host side error plumbing -----------------------------------------------------------------
int zg_set_error(const char *fmt, ...);          // returns 1
void zg_count_launch(int n = 1);
int zg_check_launch(const char *what);           // cudaPeekAtLastError -> error code

#define ZG_REQUIRE(cond, ...)                        \
    do {                                             \
        if (!(cond)) return zg_set_error(__VA_ARGS__); \
    } while (0)

static inline int zg_dtype_size(int dt) { return dt == ZG_F32 ? 4 : 2; }

// ---- device helpers ---------------------------------------------------------------------------
#define ZG_LOG2E 1.4426950408889634f
#define ZG_LN2 0.6931471805599453f

template <typename T> __device__ __forceinline__ float zg_to_float(T v);
template <> __device__ __forceinline__ float zg_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float zg_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float zg_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T zg_from_float(float v);
template <> __device__ __forceinline__ float zg_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half zg_from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 zg_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float zg_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float zg_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float zg_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softplus with the reference's threshold (selective_scan_fwd_kernel.cuh:153-156): x <= 20 ?
// log1p(exp(x)) : x.  exp via MUFU.EX2; log1p via a series for tiny e (where lg2.approx of 1+e has
// no relative accuracy) and MUFU.LG2 otherwise.  Relative error < 2e-6 over the whole range.
__device__ __forceinline__ float zg_softplus20(float x) {
    if (x > 20.f) return x;
    const float e = zg_ex2(x * ZG_LOG2E);
    if (e < 0.03125f) {
        // log1p(e) = e - e^2/2 + e^3/3 - e^4/4 + e^5/5   (|err| < e^6/6 < 2e-10 * e)
        return e * (1.f + e * (-0.5f + e * (0.33333334f + e * (-0.25f + e * 0.2f))));
    }
    return zg_lg2(1.f + e) * ZG_LN2;
}

__device__ __forceinline__ float zg_sigmoid(float x) { return zg_rcp(1.f + zg_ex2(-x * ZG_LOG2E)); }
__device__ __forceinline__ float zg_silu(float x) { return x * zg_sigmoid(x); }

// cp.async (LDGSTS) 16-byte copy global -> shared, L2 only (streamed data, no L1 allocation)
__device__ __forceinline__ void zg_cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void zg_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void zg_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float zg_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
