// Selective-scan (S6) forward for sm_100a -- sequential-in-L, parallel over (batch, channel, state).
//
// Replaces selective_scan_fwd_kernel (dis_mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303).
// The reference maps one CTA to one (batch, channel) row and runs a CUB block scan over L for each
// of the N states in turn (17.4 issued SASS instructions per (b,d,l,n) update, 3 BAR.SYNC per state
// iteration).  At the batch sizes of the sampling path there are B*E >= 80k
// independent recurrences -- more than enough to fill 148 SMs without parallelising over L -- so here
// every THREAD owns one channel, keeps its N states in registers and walks L sequentially:
// 1 FMUL + 1 MUFU.EX2 + 1 FMUL + 2 FFMA per state update, no block-wide scan, no shuffles.
// The (u, delta, z) tiles and the B/C rows of the next steps are staged into shared memory by a
// cp.async (LDGSTS) ring so HBM latency never sits on the recurrence's dependency chain.
//
// Two tile loaders behind one kernel body:
//   SEQ = true   activations (batch, dim, seqlen) with seqlen contiguous -- the reference layout of
//                selective_scan_fn; a thread reads 16-byte vectors along its own row (smem rows
//                padded to an odd number of 16-byte units -> conflict-free LDS.128).
//   SEQ = false  activations (batch, seqlen, dim) with dim contiguous -- the token-major layout of
//                the fused model path; a warp reads 32 consecutive channels of one token (coalesced),
//                B/C come straight from x_dbl rows, and the z half can be gathered through z_rowmap
//                (the zigzag permutation) so no permuted copy of xz is ever materialised.
#pragma once
#include "zg_common.cuh"

namespace zg {

constexpr int SCAN_CH = 64;      // channels (= threads) per CTA
constexpr int SCAN_TL = 16;      // time steps per pipeline stage

template <typename T, bool SEQ> struct ScanSmem {
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int NSTAGE = SEQ ? 2 : 3;
    // SEQ: row = TL elements + one 16-byte pad;  !SEQ: row = CH elements
    static constexpr int ACT_ROW_BYTES = SEQ ? (SCAN_TL * (int)sizeof(T) + 16) : (SCAN_CH * (int)sizeof(T));
    static constexpr int ACT_ROWS = SEQ ? SCAN_CH : SCAN_TL;
    static constexpr int ACT_BYTES = ACT_ROW_BYTES * ACT_ROWS;   // one tensor, one stage
    static __host__ __device__ constexpr int raw_bc_bytes(int NS) { return NS * SCAN_TL * (int)sizeof(T); }
    static __host__ __device__ constexpr int stage_bytes(int NS) { return 3 * ACT_BYTES + 2 * raw_bc_bytes(NS); }
    static __host__ __device__ constexpr int total_bytes(int NS) {
        return NSTAGE * stage_bytes(NS) + SCAN_TL * 2 * NS * (int)sizeof(float);
    }
};

// copy one 16-byte chunk (VEC elements) global -> shared; nvalid = how many leading elements exist
template <typename T>
__device__ __forceinline__ void copy_chunk(T *sdst, const T *gsrc, int nvalid) {
    constexpr int VEC = 16 / sizeof(T);
    if (nvalid >= VEC && ((reinterpret_cast<uintptr_t>(gsrc) & 15) == 0)) {
        zg_cp_async16(sdst, gsrc);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sdst[i] = (i < nvalid) ? gsrc[i] : zg_from_float<T>(0.f);
    }
}

template <typename T, int NS, bool SEQ, bool CONSTBC>
__global__ void __launch_bounds__(SCAN_CH) scan_fwd_kernel(const zg_scan_params p) {
    using SM = ScanSmem<T, SEQ>;
    constexpr int VEC = SM::VEC;
    constexpr int TL = SCAN_TL, CH = SCAN_CH, NSTAGE = SM::NSTAGE;
    extern __shared__ __align__(16) unsigned char smem[];
    float *bcf = reinterpret_cast<float *>(smem + NSTAGE * SM::stage_bytes(NS));   // [TL][2*NS]

    const int tid = threadIdx.x;
    const int E = p.dim, L = p.seqlen, N = p.dstate;
    const int per_group = E / p.ngroups;
    const int tiles_per_group = (per_group + CH - 1) / CH;
    const int tiles = tiles_per_group * p.ngroups;
    const int b = blockIdx.x / tiles;           // 1-D grid: batch may exceed 65535 (video temporal scans)
    const int tile = blockIdx.x % tiles;
    const int g = tile / tiles_per_group;
    const int e0 = g * per_group + (tile % tiles_per_group) * CH;
    const int e_end = min(e0 + CH, (g + 1) * per_group);
    const int e = e0 + tid;
    const bool active = e < e_end;
    const bool has_z = p.z != nullptr;
    const bool softplus = (p.flags & ZG_SCAN_DELTA_SOFTPLUS) != 0;
    const bool varB = (p.flags & ZG_SCAN_VARIABLE_B) != 0;
    const bool varC = (p.flags & ZG_SCAN_VARIABLE_C) != 0;

    const T *gu = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_sb;
    const T *gd = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.delta_sb;
    const T *gz = has_z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_sb : nullptr;
    const T *gB = varB ? reinterpret_cast<const T *>(p.B) + (int64_t)b * p.B_sb + (int64_t)g * p.B_sg : nullptr;
    const T *gC = varC ? reinterpret_cast<const T *>(p.C) + (int64_t)b * p.C_sb + (int64_t)g * p.C_sg : nullptr;
    T *gout = reinterpret_cast<T *>(p.out) + (int64_t)b * p.out_sb;

    // ---- per-thread constants and state --------------------------------------------------------
    float Al2[NS], h[NS];
    float Bc[CONSTBC ? NS : 1], Cc[CONSTBC ? NS : 1];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
        Al2[n] = (active && n < N) ? p.A[(int64_t)e * N + n] * ZG_LOG2E : 0.f;
        h[n] = 0.f;
        if (CONSTBC) {
            Bc[n] = (!varB && active && n < N) ? reinterpret_cast<const float *>(p.B)[(int64_t)e * N + n] : 0.f;
            Cc[n] = (!varC && active && n < N) ? reinterpret_cast<const float *>(p.C)[(int64_t)e * N + n] : 0.f;
        }
    }
    const float Dv = (p.D && active) ? p.D[e] : 0.f;
    const float bias = (p.delta_bias && active) ? p.delta_bias[e] : 0.f;

    const int nstages = (L + TL - 1) / TL;

    // ---- stage loader ---------------------------------------------------------------------------
    auto issue_stage = [&](int s) {
        if (s < nstages) {
            unsigned char *st = smem + (s % NSTAGE) * SM::stage_bytes(NS);
            const int l0 = s * TL;
            if (SEQ) {
                constexpr int CPR = TL * (int)sizeof(T) / 16;      // 16-byte chunks per row
                const int nk = has_z ? 3 : 2;
                for (int it = tid; it < nk * CH * CPR; it += CH) {
                    const int k = it / (CH * CPR), rem = it % (CH * CPR);
                    const int c = rem / CPR, j = rem % CPR;
                    const int ee = e0 + c;
                    if (ee >= e_end) continue;
                    const int l = l0 + j * VEC;
                    const int nvalid = min(L - l, VEC);
                    if (nvalid <= 0) continue;
                    const T *src = (k == 0) ? gu + (int64_t)ee * p.u_sd + l
                                 : (k == 1) ? gd + (int64_t)ee * p.delta_sd + l
                                            : gz + (int64_t)ee * p.z_sd + l;
                    copy_chunk<T>(reinterpret_cast<T *>(st + k * SM::ACT_BYTES + c * SM::ACT_ROW_BYTES + j * 16), src, nvalid);
                }
                // B / C rows: raw[n][TL]
                const int nbc = (varB ? 1 : 0) + (varC ? 1 : 0);
                for (int it = tid; it < nbc * N * CPR; it += CH) {
                    const int w = it / (N * CPR), rem = it % (N * CPR);
                    const int n = rem / CPR, j = rem % CPR;
                    const bool isB = varB && (w == 0);
                    const int l = l0 + j * VEC;
                    const int nvalid = min(L - l, VEC);
                    if (nvalid <= 0) continue;
                    const T *src = isB ? gB + (int64_t)n * p.B_sn + l : gC + (int64_t)n * p.C_sn + l;
                    T *dst = reinterpret_cast<T *>(st + 3 * SM::ACT_BYTES + (isB ? 0 : SM::raw_bc_bytes(NS))) + n * TL + j * VEC;
                    copy_chunk<T>(dst, src, nvalid);
                }
            } else {
                constexpr int CPR = CH * (int)sizeof(T) / 16;      // chunks per token row
                const int nk = has_z ? 3 : 2;
                for (int it = tid; it < nk * TL * CPR; it += CH) {
                    const int k = it / (TL * CPR), rem = it % (TL * CPR);
                    const int t = rem / CPR, j = rem % CPR;
                    const int l = l0 + t;
                    if (l >= L) continue;
                    const int ee = e0 + j * VEC;
                    const int nvalid = min(e_end - ee, VEC);
                    if (nvalid <= 0) continue;
                    const T *src;
                    if (k == 0) src = gu + (int64_t)l * p.u_sl + ee;
                    else if (k == 1) src = gd + (int64_t)l * p.delta_sl + ee;
                    else src = gz + (int64_t)(p.z_rowmap ? p.z_rowmap[l] : l) * p.z_sl + ee;
                    copy_chunk<T>(reinterpret_cast<T *>(st + k * SM::ACT_BYTES + t * SM::ACT_ROW_BYTES + j * 16), src, nvalid);
                }
                // B / C: raw[t][NS] (state contiguous)
                constexpr int BPR = (NS * (int)sizeof(T) + 15) / 16;
                const int nbc = (varB ? 1 : 0) + (varC ? 1 : 0);
                for (int it = tid; it < nbc * TL * BPR; it += CH) {
                    const int w = it / (TL * BPR), rem = it % (TL * BPR);
                    const int t = rem / BPR, j = rem % BPR;
                    const bool isB = varB && (w == 0);
                    const int l = l0 + t;
                    if (l >= L) continue;
                    const int n0 = j * VEC;
                    const int nvalid = min(N - n0, VEC);
                    if (nvalid <= 0) continue;
                    const T *src = isB ? gB + (int64_t)l * p.B_sl + n0 : gC + (int64_t)l * p.C_sl + n0;
                    T *dst = reinterpret_cast<T *>(st + 3 * SM::ACT_BYTES + (isB ? 0 : SM::raw_bc_bytes(NS))) + t * NS + n0;
                    copy_chunk<T>(dst, src, nvalid);
                }
            }
        }
        zg_cp_async_commit();
    };

    // ---- one recurrence step --------------------------------------------------------------------
    auto step = [&](int t, float uu, float dd, float zz) -> float {
        float dl = dd + bias;
        if (softplus) dl = zg_softplus20(dl);
        const float du = dl * uu;
        float y = Dv * uu;
        const float4 *bc4 = reinterpret_cast<const float4 *>(bcf + t * 2 * NS);
#pragma unroll
        for (int q = 0; q < NS / 4; ++q) {
            float4 Bv = bc4[q], Cv = bc4[NS / 4 + q];
            float Bn[4] = {Bv.x, Bv.y, Bv.z, Bv.w}, Cn[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = q * 4 + i;
                if (CONSTBC) {
                    if (!varB) Bn[i] = Bc[n];
                    if (!varC) Cn[i] = Cc[n];
                }
                const float a = zg_ex2(dl * Al2[n]);
                h[n] = fmaf(a, h[n], du * Bn[i]);
                y = fmaf(Cn[i], h[n], y);
            }
        }
        if (has_z) y *= zg_silu(zz);
        return y;
    };

    // ---- pipeline -------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) issue_stage(s);

    for (int s = 0; s < nstages; ++s) {
        issue_stage(s + NSTAGE - 1);
        zg_cp_async_wait<NSTAGE - 1>();
        __syncthreads();
        unsigned char *st = smem + (s % NSTAGE) * SM::stage_bytes(NS);
        const int l0 = s * TL;
        const int nsteps = min(TL, L - l0);
        // raw B/C (I/O dtype) -> fp32 [t][B0..B(NS-1) C0..C(NS-1)], zero padded
        {
            const T *rawB = reinterpret_cast<const T *>(st + 3 * SM::ACT_BYTES);
            const T *rawC = reinterpret_cast<const T *>(st + 3 * SM::ACT_BYTES + SM::raw_bc_bytes(NS));
            for (int it = tid; it < 2 * TL * NS; it += CH) {
                const int w = it / (TL * NS), rem = it % (TL * NS);
                int t, n;
                if (SEQ) { n = rem / TL; t = rem % TL; } else { t = rem / NS; n = rem % NS; }
                float v = 0.f;
                if (n < N && t < nsteps) {
                    if (w == 0 && varB) v = zg_to_float<T>(SEQ ? rawB[n * TL + t] : rawB[t * NS + n]);
                    if (w == 1 && varC) v = zg_to_float<T>(SEQ ? rawC[n * TL + t] : rawC[t * NS + n]);
                }
                bcf[t * 2 * NS + w * NS + n] = v;
            }
        }
        __syncthreads();

        if (active) {
            if (SEQ) {
                const unsigned char *ru = st + 0 * SM::ACT_BYTES + tid * SM::ACT_ROW_BYTES;
                const unsigned char *rd = st + 1 * SM::ACT_BYTES + tid * SM::ACT_ROW_BYTES;
                const unsigned char *rz = st + 2 * SM::ACT_BYTES + tid * SM::ACT_ROW_BYTES;
                T *orow = gout + (int64_t)e * p.out_sd + l0;
#pragma unroll
                for (int tv = 0; tv < TL / VEC; ++tv) {
                    if (tv * VEC >= nsteps) break;
                    union { uint4 v; T e[VEC]; } U, Dl, Z, O;
                    U.v = *reinterpret_cast<const uint4 *>(ru + tv * 16);
                    Dl.v = *reinterpret_cast<const uint4 *>(rd + tv * 16);
                    if (has_z) Z.v = *reinterpret_cast<const uint4 *>(rz + tv * 16);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const int t = tv * VEC + i;
                        float y = 0.f;
                        if (t < nsteps)
                            y = step(t, zg_to_float<T>(U.e[i]), zg_to_float<T>(Dl.e[i]), has_z ? zg_to_float<T>(Z.e[i]) : 0.f);
                        O.e[i] = zg_from_float<T>(y);
                    }
                    T *dst = orow + tv * VEC;
                    if (tv * VEC + VEC <= nsteps && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                        *reinterpret_cast<uint4 *>(dst) = O.v;
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i)
                            if (tv * VEC + i < nsteps) dst[i] = O.e[i];
                    }
                }
            } else {
                const T *su = reinterpret_cast<const T *>(st + 0 * SM::ACT_BYTES) + tid;
                const T *sd = reinterpret_cast<const T *>(st + 1 * SM::ACT_BYTES) + tid;
                const T *sz = reinterpret_cast<const T *>(st + 2 * SM::ACT_BYTES) + tid;
                T *ocol = gout + (int64_t)l0 * p.out_sl + e;
                if (nsteps == TL) {
#pragma unroll 4
                    for (int t = 0; t < TL; ++t) {
                        const float y = step(t, zg_to_float<T>(su[t * CH]), zg_to_float<T>(sd[t * CH]),
                                             has_z ? zg_to_float<T>(sz[t * CH]) : 0.f);
                        ocol[(int64_t)t * p.out_sl] = zg_from_float<T>(y);
                    }
                } else {
                    for (int t = 0; t < nsteps; ++t) {
                        const float y = step(t, zg_to_float<T>(su[t * CH]), zg_to_float<T>(sd[t * CH]),
                                             has_z ? zg_to_float<T>(sz[t * CH]) : 0.f);
                        ocol[(int64_t)t * p.out_sl] = zg_from_float<T>(y);
                    }
                }
            }
            // recompute seeds for the backward pass
            if (p.ckpt) {
                const int lend = l0 + nsteps;
                if (lend % p.ckpt_every == 0 || lend == L) {
                    const int k = (lend - 1) / p.ckpt_every;
                    const int nck = (L + p.ckpt_every - 1) / p.ckpt_every;
                    float *dst = p.ckpt + (((int64_t)b * E + e) * nck + k) * N;
#pragma unroll
                    for (int n = 0; n < NS; ++n)
                        if (n < N) dst[n] = h[n];
                }
            }
        }
        __syncthreads();   // stage buffer and bcf are recycled by the next iteration
    }

    if (active && p.last_state) {
        float *dst = p.last_state + ((int64_t)b * E + e) * N;
#pragma unroll
        for (int n = 0; n < NS; ++n)
            if (n < N) dst[n] = h[n];
    }
}

template <typename T, int NS, bool SEQ, bool CONSTBC>
int launch_scan_fwd(const zg_scan_params &p, cudaStream_t stream) {
    using SM = ScanSmem<T, SEQ>;
    const int per_group = p.dim / p.ngroups;
    const int tiles = p.ngroups * ((per_group + SCAN_CH - 1) / SCAN_CH);
    const int smem = SM::total_bytes(NS);
    auto kern = scan_fwd_kernel<T, NS, SEQ, CONSTBC>;
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err != cudaSuccess) return zg_set_error("scan_fwd: cudaFuncSetAttribute(%d B smem): %s", smem, cudaGetErrorString(err));
        attr_set = true;
    }
    const long long nblk = (long long)tiles * p.batch;
    if (nblk > 0x7fffffffLL) return zg_set_error("scan_fwd: grid too large (%lld CTAs)", nblk);
    if (nblk == 0) return 0;
    kern<<<(unsigned)nblk, SCAN_CH, smem, stream>>>(p);
    zg_count_launch();
    return zg_check_launch("scan_fwd");
}

// one translation unit per I/O dtype (parallel compilation)
template <typename T> int dispatch_scan_fwd(const zg_scan_params &p, bool seq, bool constbc, cudaStream_t stream) {
    const int N = p.dstate;
#define ZG_SCAN_CASE(NSV)                                                                   \
    if (N <= NSV) {                                                                         \
        if (seq) return launch_scan_fwd<T, NSV, true, false>(p, stream);                    \
        return launch_scan_fwd<T, NSV, false, false>(p, stream);                            \
    }
    if (constbc) {
        if (N <= 8) return seq ? launch_scan_fwd<T, 8, true, true>(p, stream) : launch_scan_fwd<T, 8, false, true>(p, stream);
        if (N <= 16) return seq ? launch_scan_fwd<T, 16, true, true>(p, stream) : launch_scan_fwd<T, 16, false, true>(p, stream);
        return zg_set_error("selective_scan_fwd: constant (non input-dependent) B/C supports dstate <= 16, got %d", N);
    }
    ZG_SCAN_CASE(8)
    ZG_SCAN_CASE(16)
    ZG_SCAN_CASE(32)
    ZG_SCAN_CASE(64)
#undef ZG_SCAN_CASE
    return zg_set_error("selective_scan_fwd: dstate %d > 64 not supported", N);
}

int scan_fwd_f32(const zg_scan_params &p, bool seq, bool constbc, cudaStream_t s);
int scan_fwd_f16(const zg_scan_params &p, bool seq, bool constbc, cudaStream_t s);
int scan_fwd_bf16(const zg_scan_params &p, bool seq, bool constbc, cudaStream_t s);

}  // namespace zg
