// K1 (DMMA variant) -- placeholder until the tensor-core kernel lands; the host falls back to the
// generic variant when this reports "unsupported".
#include "dto_internal.h"

bool bilinear_dmma_supported(int n, int m) { return false; }

bool launch_bilinear_dmma(const DProb& P, int ii, const double* Z, const double* mu, double* g, double* jac, EvalFlags f,
                          cudaStream_t st, long long* launches) {
    return false;
}
