// forward-pass instantiations: UNGM (1-D state, 1-D measurement)
// small bodies: libm inlined (the out-of-line copies exist for the 5-D bodies, which overflow the instruction cache)
#define SSM_INLINE_MATH 1
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_ungm(const FilterLaunch &L) { return dispatch_filter_model<DynUngm, ObsUngm<1, 0>, 128, 4>(L); }
}  // namespace ssm
