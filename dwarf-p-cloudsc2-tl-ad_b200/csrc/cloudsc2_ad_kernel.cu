// TEMPORARY placeholder: adjoint launchers (replaced by the real kernel).
#include "cloudsc2_launch.h"
cudaError_t csc2_launch_ad(const KConst &, const Geom &, const TrajIn &, const TrajOut &,
                           const IncIn &, const IncOut &, const ADOpts &, cudaStream_t) {
  return cudaErrorNotSupported;
}
cudaError_t csc2_launch_ad_finalize(const Geom &, const double *, const double *, double *, double *,
                                    cudaStream_t) {
  return cudaErrorNotSupported;
}
