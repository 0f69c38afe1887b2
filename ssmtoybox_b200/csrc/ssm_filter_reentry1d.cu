// forward-pass instantiations: vertically falling reentry body (3-D state) + range measurement
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_reentry1d(const FilterLaunch &L) { return dispatch_filter_model<DynReentry1D, ObsRange<3, 0>, 128, 4>(L); }
}  // namespace ssm
