// selective-scan forward, I/O dtype __half (one TU per dtype so the instantiations compile in parallel)
#include "scan_fwd.cuh"
#include "scan_fwd_tpc2.cuh"
#include "scan_fwd_tma.cuh"
namespace zg {
int scan_fwd_f16(const zg_scan_params &p, bool seq, bool constbc, cudaStream_t s) {
    return dispatch_scan_fwd<__half>(p, seq, constbc, s);
}
}  // namespace zg
