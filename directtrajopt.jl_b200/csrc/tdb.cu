// K7 -- TimeDependentBilinearIntegrator interval kernel (placeholder; see tdb_available()).
#include "dto_internal.h"

bool tdb_available() { return false; }

void launch_tdb(const DProb& P, int ii, const double* Z, const double* mu, double* g, double* jac, EvalFlags f, cudaStream_t st,
                long long* launches) {}
