// selective-scan forward, I/O dtype float (one TU per dtype so the instantiations compile in parallel)
#include "scan_fwd.cuh"
namespace zg {
int scan_fwd_f32(const zg_scan_params &p, bool seq, bool constbc, cudaStream_t s) {
    return dispatch_scan_fwd<float>(p, seq, constbc, s);
}
}  // namespace zg
