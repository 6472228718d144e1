// selective-scan forward, warp-private pipeline with mixed 32- / 16-channel warps, I/O dtype __half (own TU)
#include "scan_fwd_wph.cuh"
namespace zg {
int scan_fwd_wph_f16(const zg_scan_params &p, cudaStream_t stream, int nd, int ns) { return wph_launch_variant<__half>(p, stream, nd, ns); }
}  // namespace zg
