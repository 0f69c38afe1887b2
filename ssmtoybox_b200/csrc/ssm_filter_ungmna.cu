// forward-pass instantiations: UNGM with NON-additive process and measurement noise (augmented-state transforms)
// small bodies: libm inlined (the out-of-line copies exist for the 5-D bodies, which overflow the instruction cache)
#define SSM_INLINE_MATH 1
#include "ssm_filter_dispatch.cuh"
namespace ssm {
int filter_ungmna(const FilterLaunch &L) { return dispatch_filter_model<DynUngmNA, ObsUngmNA<1, 0>, 128, 4>(L); }
}  // namespace ssm
