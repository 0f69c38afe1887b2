// forward-pass instantiations: pendulum (2-D state, 1-D measurement)
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_pendulum(const FilterLaunch &L) { return dispatch_filter_model<DynPendulum, ObsPendulum<2, 0>, 128, 4>(L); }
}  // namespace ssm
