// tc_scan.cu — tensor-core scan path (placeholder until the tcgen05 kernels land).
#include "tc_scan.cuh"
#include "../../include/fvdb.h"

namespace fvdb {

bool tc_supported(uint32_t) { return false; }

int tc_ivf_search(TcScratch&, const TcSearchArgs&, cudaStream_t, size_t*, uint32_t*, uint32_t*,
                  std::string* err) {
    if (err) *err = "tensor-core scan not built";
    return FVDB_ERR_INVALID_CONFIG;
}
int tc_flat_search(TcScratch&, const TcFlatArgs&, cudaStream_t, size_t*, uint32_t*, uint32_t*,
                   std::string* err) {
    if (err) *err = "tensor-core scan not built";
    return FVDB_ERR_INVALID_CONFIG;
}
int tc_assign(TcScratch&, const float*, uint32_t, const float*, uint64_t, uint32_t, uint64_t*,
              cudaStream_t, size_t*, std::string* err) {
    if (err) *err = "tensor-core scan not built";
    return FVDB_ERR_INVALID_CONFIG;
}
void tc_release(TcScratch&) {}

}  // namespace fvdb
