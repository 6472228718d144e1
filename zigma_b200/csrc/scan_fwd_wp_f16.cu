// selective-scan forward, warp-private pipeline, I/O dtype __half (own TU: compiles in parallel with the other scan kernels)
#include "scan_fwd_wp.cuh"
namespace zg {
int scan_fwd_wp_f16(const zg_scan_params &p, cudaStream_t stream, int mode) { return wp_launch_variant<__half>(p, stream, mode); }
}  // namespace zg
