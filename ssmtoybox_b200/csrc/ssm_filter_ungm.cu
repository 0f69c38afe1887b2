// forward-pass instantiations: UNGM (1-D state, 1-D measurement)
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_ungm(const FilterLaunch &L) { return dispatch_filter_model<DynUngm, ObsUngm<1, 0>, 128, 4>(L); }
}  // namespace ssm
