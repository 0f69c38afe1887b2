// forward-pass instantiations: constant velocity (4-D state) + radar on state_index = [0, 1] (the default) or [0, 2]
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_constvel01(const FilterLaunch &L) { return dispatch_filter_model<DynConstVel, ObsRadar<4, 0, 1>, 128, 4>(L); }
int filter_constvel02(const FilterLaunch &L) { return dispatch_filter_model<DynConstVel, ObsRadar<4, 0, 2>, 128, 4>(L); }
}  // namespace ssm
