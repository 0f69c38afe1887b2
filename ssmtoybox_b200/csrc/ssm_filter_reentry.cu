// forward-pass instantiations: reentry vehicle (5-D state) + radar on the leading two components
#define SSM_PAIR_MODEL 1
#include "ssm_filter_dispatch.cuh"
#ifndef SSM_MINB_5D
#define SSM_MINB_5D 3
#endif
#ifndef SSM_THREADS_5D
#define SSM_THREADS_5D 128
#endif
namespace ssm {
int filter_reentry(const FilterLaunch &L) { return dispatch_filter_model<DynReentry, ObsRadar<5, 0, 1>, SSM_THREADS_5D, SSM_MINB_5D>(L); }
}  // namespace ssm
