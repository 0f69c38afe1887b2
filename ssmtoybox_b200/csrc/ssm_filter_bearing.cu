// forward-pass instantiations: coordinated turn (5-D state) + bearings from 4 sensors on state_index = [0, 2]
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_coordturn_bearing(const FilterLaunch &L) { return dispatch_filter_model<DynCoordTurn, ObsBearing4<5, 0, 2>, 128, 3>(L); }
}  // namespace ssm
