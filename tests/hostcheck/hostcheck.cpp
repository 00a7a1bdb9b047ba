// TEST-ONLY: compiles the host/device "core" headers of the CUDA kernels with g++ so that their
// algebra can be checked against the oracle on a machine without a GPU.  Never loaded by the product.
#include <cstdint>

#include "mc_core.cuh"

extern "C" {

void hostcheck_mc(const mc_params_in* prm, const double* deps, const double* sigma_n, double* C_tang, double* sigma,
                  int32_t* niter, double* yielding, double* norm_res, double* dlambda, int64_t n) {
  mc_consts k;
  mc_make_consts(*prm, k);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i)
    mc_point(k, deps + 4 * i, sigma_n + 4 * i, C_tang + 16 * i, sigma + 4 * i, niter[i], yielding[i], norm_res[i],
             dlambda[i]);
}
}
