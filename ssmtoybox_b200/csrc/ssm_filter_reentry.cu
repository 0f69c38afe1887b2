// forward-pass instantiations: reentry vehicle (5-D state) + radar on the leading two components
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_reentry(const FilterLaunch &L) { return dispatch_filter_model<DynReentry, ObsRadar<5, 0, 1>, 128, 2>(L); }
}  // namespace ssm
