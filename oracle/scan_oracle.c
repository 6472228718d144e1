/* TEST INFRASTRUCTURE ONLY -- plain C restatement of the S6 selective-scan forward, used as the
 * CPU checker for sizes the torch restatement (oracle/zigma_oracle.py) is too slow for and as the
 * `cpu_baseline` / `--impl reference` leg of bench.py.  Never linked into the product library.
 *
 * Follows dis_mamba/mamba_ssm/ops/selective_scan_interface.py:86-152 (selective_scan_ref) with the
 * fp32 semantics of dis_mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:153-171,216-261,
 * 280-298: delta' = softplus(delta + bias) (identity above 20), h <- exp(delta' A) h + delta' B u,
 * y = C.h + D u, out = y * silu(z).  Pinned against the reference by oracle/gen_golden.py
 * (tests/golden/scan_*.npz) and tests/test_oracle_golden.py.
 *
 * Layout: u, delta, z, out (Bt, E, L) row-major; A (E, N); B, C (Bt, G, N, L); D, delta_bias (E).
 * Channel e uses group e / (E / G).  last_state (Bt, E, N) may be NULL.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

int zigma_oracle_scan_fwd(const float *u, const float *delta, const float *A, const float *B,
                          const float *C, const float *D, const float *z, const float *delta_bias,
                          int delta_softplus, float *out, float *last_state,
                          int Bt, int E, int L, int N, int G)
{
    if (N > 256 || G <= 0 || E % G != 0) return 1;
    const int per_group = E / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < Bt; ++b) {
        for (int e = 0; e < E; ++e) {
            float h[256];
            for (int n = 0; n < N; ++n) h[n] = 0.0f;
            const size_t row = ((size_t)b * E + e) * (size_t)L;
            const float *Bg = B + ((size_t)b * G + e / per_group) * (size_t)N * L;
            const float *Cg = C + ((size_t)b * G + e / per_group) * (size_t)N * L;
            const float bias = delta_bias ? delta_bias[e] : 0.0f;
            const float Dv = D ? D[e] : 0.0f;
            for (int l = 0; l < L; ++l) {
                float d = delta[row + l] + bias;
                if (delta_softplus && d <= 20.0f) d = log1pf(expf(d));
                const float uu = u[row + l];
                const float du = d * uu;
                float y = 0.0f;
                for (int n = 0; n < N; ++n) {
                    h[n] = expf(d * A[(size_t)e * N + n]) * h[n] + du * Bg[(size_t)n * L + l];
                    y += Cg[(size_t)n * L + l] * h[n];
                }
                y += Dv * uu;
                if (z) {
                    const float zz = z[row + l];
                    y *= zz / (1.0f + expf(-zz));
                }
                out[row + l] = y;
            }
            if (last_state)
                for (int n = 0; n < N; ++n) last_state[((size_t)b * E + e) * N + n] = h[n];
        }
    }
    return 0;
}

/* Depthwise causal conv1d + optional SiLU, (Bt, E, L) layout, weight (E, W).
 * dis_causal_conv1d/causal_conv1d/causal_conv1d_interface.py:49-65. */
int zigma_oracle_conv1d_fwd(const float *x, const float *w, const float *bias, int silu, float *out,
                            int Bt, int E, int L, int W)
{
    if (W < 1 || W > 8) return 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < Bt; ++b) {
        for (int e = 0; e < E; ++e) {
            const size_t row = ((size_t)b * E + e) * (size_t)L;
            for (int l = 0; l < L; ++l) {
                float acc = bias ? bias[e] : 0.0f;
                for (int k = 0; k < W; ++k) {
                    const int src = l - (W - 1 - k);
                    if (src >= 0) acc += w[(size_t)e * W + k] * x[row + src];
                }
                if (silu) acc = acc / (1.0f + expf(-acc));
                out[row + l] = acc;
            }
        }
    }
    return 0;
}

/* number of OpenMP threads of the two loops above (0 = leave the runtime default).  bench.py's CPU arm calls this: launchers such
 * as torchrun export OMP_NUM_THREADS=1 into every rank, which would silently turn the all-cores baseline into a 1-thread run. */
#ifdef _OPENMP
#include <omp.h>
int zigma_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int zigma_oracle_set_threads(int n) { (void)n; return 1; }
#endif
