// selective-scan forward, warp-private pipeline, I/O dtype __nv_bfloat16 (own TU: compiles in parallel with the other scan kernels)
#include "scan_fwd_wp.cuh"
namespace zg {
int scan_fwd_wp_bf16(const zg_scan_params &p, cudaStream_t stream, int mode) { return wp_launch_variant<__nv_bfloat16>(p, stream, mode); }
}  // namespace zg
