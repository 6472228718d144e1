// Selective-scan backward, dstate == 16 specialisation for sm_100a: FOUR threads per channel, four
// states each ("quad" = lanes 4j..4j+3 of a warp), 64 channels x one batch row per 256-thread CTA.
// Same math as the generic kernel in scan_bwd.cu (reference: selective_scan_bwd_kernel.cuh:186-213,
// 252-296,439-452); what changes is the mapping, chosen for occupancy and memory behaviour on B200:
//
//   * parked states (h_{l-1} of the 8 steps of a chunk, recomputed from the forward's checkpoint)
//     cost 64 B per (channel, step) whatever the thread mapping -- with 4 threads per channel that is
//     128 B per THREAD, so 4 CTAs = 32 warps fit an SM (the one-thread-per-channel kernel: 6 warps);
//     one LDS.128 / STS.128 per thread-step, conflict free ([step][channel][quad] float4);
//   * every global access goes through a per-chunk cooperative prologue / epilogue, so it is coalesced
//     for BOTH layouts the ABI allows (channel-first (b, d, l) as the reference's autograd passes,
//     token-major (b l, d) as the engine uses) and each softplus / sigmoid is evaluated ONCE per
//     (channel, step), not once per thread of the quad;
//   * dB / dC (sums over the channels that share a group): 7-shuffle transpose-reduce over the 8
//     channels of a warp, cross-warp sum through shared memory, then one fp32 atomic per
//     (CTA, state, step) with 8 consecutive steps per 32-byte sector;
//   * packed FFMA2 math on state pairs, as in the forward;
//   * z_rowmap (the forward's fused permutation of the gate) redirects the z reads and the dz writes.
#pragma once
#include "zg_common.cuh"

namespace zg {

constexpr int Q4_CH = 64;          // channels per CTA
constexpr int Q4_THREADS = 256;
constexpr int Q4_TS = 8;           // steps per chunk == ckpt_every of the forward
constexpr int Q4_RED_LD = Q4_TS + 1;

struct Q4Smem {
    static constexpr int HS = Q4_TS * Q4_CH * 4 * 16;              // float4 [TS][CH][4]
    static constexpr int SC = Q4_TS * Q4_CH * 16;                  // float4 [TS][CH]  (delta', u, dy, gz)
    static constexpr int BC = Q4_TS * 32 * 4;                      // float  [TS][B0..15 C0..15]
    static constexpr int RED = 8 * 32 * Q4_RED_LD * 4;             // float  [warp][slot][TS+1]
    static constexpr int OUT = 3 * Q4_TS * Q4_CH * 4;              // float  [du|ddelta|dz][TS][CH]
    static constexpr int TOTAL = HS + SC + BC + RED + OUT;
};

template <typename T>
__global__ void __launch_bounds__(Q4_THREADS, 4) scan_bwd_q4_kernel(const zg_scan_bwd_params q) {
    const zg_scan_params &p = q.fwd;
    constexpr int TS = Q4_TS, CH = Q4_CH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *hs = reinterpret_cast<float4 *>(smem_raw);
    float4 *sc = reinterpret_cast<float4 *>(smem_raw + Q4Smem::HS);
    float *bcf = reinterpret_cast<float *>(smem_raw + Q4Smem::HS + Q4Smem::SC);
    float *red = bcf + TS * 32;
    float *outt = red + 8 * 32 * Q4_RED_LD;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = tid >> 2, qd = tid & 3;                   // channel within the CTA, quarter of the state vector
    const int E = p.dim, L = p.seqlen;
    const int per_group = E / p.ngroups;
    const int tiles_per_group = (per_group + CH - 1) / CH;
    const int tiles = tiles_per_group * p.ngroups;
    const int b = blockIdx.x / tiles;
    const int tile = blockIdx.x % tiles;
    const int g = tile / tiles_per_group;
    const int e0 = g * per_group + (tile % tiles_per_group) * CH;
    const int e_end = min(e0 + CH, (g + 1) * per_group);
    const bool active = e0 + c < e_end;
    const int e = active ? e0 + c : e0;
    const bool has_z = p.z != nullptr;
    const bool softplus = (p.flags & ZG_SCAN_DELTA_SOFTPLUS) != 0;

    // ---- cooperative (prologue / epilogue) item mapping: 64 channels x 8 steps = 2 items per thread ----
    // token-major tensors: consecutive threads -> consecutive channels; channel-first: consecutive steps
    const bool tok = (p.u_sd == 1);
    int ic[2], it[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid + j * Q4_THREADS;
        ic[j] = tok ? (i & (CH - 1)) : (i >> 3);
        it[j] = tok ? (i >> 6) : (i & 7);
    }
    float ibias[2];
    bool ich_ok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        ich_ok[j] = e0 + ic[j] < e_end;
        ibias[j] = (p.delta_bias && ich_ok[j]) ? p.delta_bias[e0 + ic[j]] : 0.f;
    }
    const T *gu = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_sb + (int64_t)e0 * p.u_sd;
    const T *gd = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.delta_sb + (int64_t)e0 * p.delta_sd;
    const T *gz = has_z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_sb + (int64_t)e0 * p.z_sd : nullptr;
    const T *gdo = reinterpret_cast<const T *>(q.dout) + (int64_t)b * q.dout_sb + (int64_t)e0 * q.dout_sd;
    const T *gB = reinterpret_cast<const T *>(p.B) + (int64_t)b * p.B_sb + (int64_t)g * p.B_sg;
    const T *gC = reinterpret_cast<const T *>(p.C) + (int64_t)b * p.C_sb + (int64_t)g * p.C_sg;
    T *gdu = reinterpret_cast<T *>(q.du) + (int64_t)b * q.du_sb + (int64_t)e0 * q.du_sd;
    T *gdd = reinterpret_cast<T *>(q.ddelta) + (int64_t)b * q.ddelta_sb + (int64_t)e0 * q.ddelta_sd;
    T *gdz = has_z ? reinterpret_cast<T *>(q.dz) + (int64_t)b * q.dz_sb + (int64_t)e0 * q.dz_sd : nullptr;
    float *gdB = q.dB + ((int64_t)b * p.ngroups + g) * (int64_t)16 * L;      // (batch, groups, dstate, seqlen)
    float *gdC = q.dC + ((int64_t)b * p.ngroups + g) * (int64_t)16 * L;
    const int nck = (L + TS - 1) / TS;
    const float4 *ck = reinterpret_cast<const float4 *>(p.ckpt + ((int64_t)b * nck * E + e) * (int64_t)16) + qd;   // (batch, n_ckpt, dim, 16)
    // B / C staging: 2 x 8 steps x 16 states = 256 items, one per thread
    const int bw = tid >> 7, brem = tid & 127;
    const bool bc_tok = (bw ? p.C_sn : p.B_sn) == 1;
    const int bn = bc_tok ? (brem & 15) : (brem >> 3), bt = bc_tok ? (brem >> 4) : (brem & 7);
    const T *gbc = bw ? gC + (int64_t)bn * p.C_sn : gB + (int64_t)bn * p.B_sn;
    const int64_t bc_sl = bw ? p.C_sl : p.B_sl;

    // ---- per-thread state: 4 states as two packed pairs --------------------------------------------------
    zg_f2 A2[2], dA[2], carry[2];
    {
        const float4 a4 = *reinterpret_cast<const float4 *>(p.A + (int64_t)e * 16 + 4 * qd);
        A2[0] = make_float2(a4.x * ZG_LOG2E, a4.y * ZG_LOG2E);
        A2[1] = make_float2(a4.z * ZG_LOG2E, a4.w * ZG_LOG2E);
    }
    dA[0] = dA[1] = carry[0] = carry[1] = zg_splat2(0.f);
    const float Dv = p.D ? p.D[e] : 0.f;
    float dD_acc = 0.f, dbias_acc = 0.f;                    // meaningful in the qd == 0 thread

    for (int k = nck - 1; k >= 0; --k) {
        const int l0 = k * TS;
        // ---- prologue: inputs -> per-(step, channel) scalars in shared memory ------------------------------
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int l = l0 + it[j];
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ich_ok[j] && l < L) {
                float d = zg_to_float<T>(gd[(int64_t)ic[j] * p.delta_sd + (int64_t)l * p.delta_sl]) + ibias[j];
                if (softplus) d = zg_softplus20(d);
                const float dout = zg_to_float<T>(gdo[(int64_t)ic[j] * q.dout_sd + (int64_t)l * q.dout_sl]);
                s.x = d;
                s.y = zg_to_float<T>(gu[(int64_t)ic[j] * p.u_sd + (int64_t)l * p.u_sl]);
                s.z = dout;
                if (has_z) {
                    const int64_t lz = p.z_rowmap ? p.z_rowmap[l] : l;       // z is read (and dz written) in token order
                    const float zz = zg_to_float<T>(gz[(int64_t)ic[j] * p.z_sd + lz * p.z_sl]);
                    const float sg = zg_sigmoid(zz);
                    s.z = dout * zz * sg;
                    s.w = dout * sg * (1.f + zz * (1.f - sg));
                }
            }
            sc[it[j] * CH + ic[j]] = s;
        }
        {
            const int l = l0 + bt;
            bcf[bt * 32 + bw * 16 + bn] = (l < L) ? zg_to_float<T>(gbc[(int64_t)l * bc_sl]) : 0.f;
        }
        __syncthreads();

        // ---- forward recompute from the checkpoint before the chunk, parking h_{l-1} -----------------------
        zg_f2 h[2];
        if (k > 0) {
            const float4 h4 = ck[(int64_t)(k - 1) * E * 4];
            h[0] = make_float2(h4.x, h4.y); h[1] = make_float2(h4.z, h4.w);
        } else {
            h[0] = h[1] = zg_splat2(0.f);
        }
#pragma unroll 1
        for (int t = 0; t < TS; ++t) {
            const float4 s = sc[t * CH + c];
            const float4 B4 = *reinterpret_cast<const float4 *>(bcf + t * 32 + 4 * qd);
            hs[(t * CH + c) * 4 + qd] = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
            const zg_f2 d2 = zg_splat2(s.x), ddu2 = zg_splat2(s.x * s.y);
            h[0] = zg_fma2(zg_ex2_mufu2(zg_mul2(d2, A2[0])), h[0], zg_mul2(ddu2, make_float2(B4.x, B4.y)));
            h[1] = zg_fma2(zg_ex2_mufu2(zg_mul2(d2, A2[1])), h[1], zg_mul2(ddu2, make_float2(B4.z, B4.w)));
        }

        // ---- reverse sweep -------------------------------------------------------------------------------------
#pragma unroll 1     // (full unrolling spills at the 64-register budget of 4 CTAs / SM)
        for (int t = TS - 1; t >= 0; --t) {
            const float4 s = sc[t * CH + c];                      // delta', u, dy, gz
            const float4 B4 = *reinterpret_cast<const float4 *>(bcf + t * 32 + 4 * qd);
            const float4 C4 = *reinterpret_cast<const float4 *>(bcf + t * 32 + 16 + 4 * qd);
            const float4 hp4 = hs[(t * CH + c) * 4 + qd];
            const zg_f2 Bp[2] = {make_float2(B4.x, B4.y), make_float2(B4.z, B4.w)};
            const zg_f2 Cp[2] = {make_float2(C4.x, C4.y), make_float2(C4.z, C4.w)};
            const zg_f2 hp[2] = {make_float2(hp4.x, hp4.y), make_float2(hp4.z, hp4.w)};
            const zg_f2 d2 = zg_splat2(s.x), ddu2 = zg_splat2(s.x * s.y), dy2 = zg_splat2(s.z);
            zg_f2 y2 = zg_splat2(0.f), s1 = zg_splat2(0.f), s2 = zg_splat2(0.f);
            float v[8];                                           // dB[4qd..4qd+3] | dC[4qd..4qd+3] of this channel
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const zg_f2 a = zg_ex2_mufu2(zg_mul2(d2, A2[i]));
                const zg_f2 hl = zg_fma2(a, hp[i], zg_mul2(ddu2, Bp[i]));
                y2 = zg_fma2(Cp[i], hl, y2);
                const zg_f2 dh = zg_fma2(dy2, Cp[i], carry[i]);   // dh_l = dy_l C_l + a_{l+1} dh_{l+1}
                carry[i] = zg_mul2(a, dh);
                const zg_f2 t1 = zg_mul2(carry[i], hp[i]);        // dh_l a_l h_{l-1}
                s1 = zg_fma2(dh, Bp[i], s1);
                s2 = zg_fma2(t1, A2[i], s2);
                dA[i] = zg_fma2(t1, d2, dA[i]);
                const zg_f2 vb = zg_mul2(dh, ddu2), vc = zg_mul2(dy2, hl);
                v[2 * i] = vb.x; v[2 * i + 1] = vb.y; v[4 + 2 * i] = vc.x; v[4 + 2 * i + 1] = vc.y;
            }
            float yq = y2.x + y2.y, s1q = s1.x + s1.y, s2q = s2.x + s2.y;
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                yq += __shfl_xor_sync(0xffffffffu, yq, o);
                s1q += __shfl_xor_sync(0xffffffffu, s1q, o);
                s2q += __shfl_xor_sync(0xffffffffu, s2q, o);
            }
            if (qd == 0) {
                const float du_ = fmaf(s.x, s1q, s.z * Dv);
                float dd = fmaf(s2q, ZG_LN2, s.y * s1q);
                if (softplus) dd *= 1.f - zg_ex2(-s.x * ZG_LOG2E);   // sigmoid(delta~) = 1 - exp(-softplus(delta~))
                dD_acc = fmaf(s.z, s.y, dD_acc);
                dbias_acc += dd;
                outt[t * CH + c] = du_;
                outt[TS * CH + t * CH + c] = dd;
                outt[2 * TS * CH + t * CH + c] = s.w * fmaf(Dv, s.y, yq);
            }
            // transpose-reduce the 8 values over the 8 channels of the warp (lane bits 4, 3, 2): afterwards the
            // lane of channel j holds the total of value j
#pragma unroll
            for (int half = 4; half >= 1; half >>= 1) {
                const bool up = (lane & (half << 2)) != 0;
#pragma unroll
                for (int j = 0; j < half; ++j) {
                    const float send = up ? v[j] : v[j + half];
                    const float keep = up ? v[j + half] : v[j];
                    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half << 2);
                }
            }
            const int vi = lane >> 2;
            const int slot = ((vi & 4) << 2) + 4 * qd + (vi & 3);       // [dB0..15 | dC0..15]
            red[(warp * 32 + slot) * Q4_RED_LD + t] = v[0];
        }
        __syncthreads();

        // ---- epilogue: coalesced stores + one atomic per (state, step) ----------------------------------------
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int l = l0 + it[j];
            if (ich_ok[j] && l < L) {
                const int o = it[j] * CH + ic[j];
                gdu[(int64_t)ic[j] * q.du_sd + (int64_t)l * q.du_sl] = zg_from_float<T>(outt[o]);
                gdd[(int64_t)ic[j] * q.ddelta_sd + (int64_t)l * q.ddelta_sl] = zg_from_float<T>(outt[TS * CH + o]);
                if (has_z) {
                    const int64_t lz = p.z_rowmap ? p.z_rowmap[l] : l;
                    gdz[(int64_t)ic[j] * q.dz_sd + lz * q.dz_sl] = zg_from_float<T>(outt[2 * TS * CH + o]);
                }
            }
        }
        {
            const int slot = tid >> 3, t = tid & 7, l = l0 + t;
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) acc += red[(w * 32 + slot) * Q4_RED_LD + t];
            if (l < L) atomicAdd((slot < 16 ? gdB : gdC) + (int64_t)(slot & 15) * L + l, acc);
        }
        // the next chunk's prologue only writes sc / bcf, which nobody reads any more; its first barrier orders
        // this epilogue's reads of outt / red before the next reverse sweep's writes
    }

    if (active) {
        float *dAe = q.dA + (int64_t)e * 16 + 4 * qd;
        atomicAdd(dAe + 0, dA[0].x); atomicAdd(dAe + 1, dA[0].y);
        atomicAdd(dAe + 2, dA[1].x); atomicAdd(dAe + 3, dA[1].y);
        if (qd == 0) {
            if (q.dD) atomicAdd(q.dD + e, dD_acc);
            if (q.ddelta_bias) atomicAdd(q.ddelta_bias + e, dbias_acc);
        }
    }
}

// returns -1 when the call does not fit this specialisation
template <typename T> int try_launch_scan_bwd_q4(const zg_scan_bwd_params &q, cudaStream_t s) {
    static int enabled = -1;
    if (enabled < 0) { const char *e = getenv("ZG_SCAN_BWD_Q4"); enabled = e ? atoi(e) : 1; }
    const zg_scan_params &p = q.fwd;
    if (!enabled || p.dstate != 16 || p.ckpt_every != Q4_TS) return -1;
    if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.ckpt)) % 16 != 0) return -1;
    const int per_group = p.dim / p.ngroups;
    const long long nblk = (long long)p.ngroups * ((per_group + Q4_CH - 1) / Q4_CH) * p.batch;
    if (nblk > 0x7fffffffLL) return -1;
    auto kern = scan_bwd_q4_kernel<T>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Q4Smem::TOTAL);
        if (err != cudaSuccess) return zg_set_error("scan_bwd(q4): cudaFuncSetAttribute(%d B smem): %s", Q4Smem::TOTAL, cudaGetErrorString(err));
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    kern<<<(unsigned)nblk, Q4_THREADS, Q4Smem::TOTAL, s>>>(q);
    zg_count_launch();
    return zg_check_launch("scan_bwd(q4)");
}

}  // namespace zg
