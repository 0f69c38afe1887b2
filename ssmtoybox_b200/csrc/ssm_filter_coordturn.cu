// forward-pass instantiations: coordinated turn (5-D state) + radar on state_index = [0, 2]
#define SSM_PAIR_MODEL 1
#include "ssm_filter_dispatch.cuh"
#ifndef SSM_MINB_5D
#define SSM_MINB_5D 3
#endif
namespace ssm {
int filter_coordturn(const FilterLaunch &L) { return dispatch_filter_model<DynCoordTurn, ObsRadar<5, 0, 2>, 128, SSM_MINB_5D>(L); }
}  // namespace ssm
