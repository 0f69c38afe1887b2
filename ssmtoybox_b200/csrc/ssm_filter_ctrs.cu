// forward-pass instantiations: constant turn rate and speed (5-D state, NON-additive 2-D noise) + radar
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_ctrs(const FilterLaunch &L) { return dispatch_filter_model<DynCtrs, ObsRadar<5, 0, 1>, 128, 3>(L); }
}  // namespace ssm
