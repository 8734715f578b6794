// E half-step kernel instantiations + dispatch for ragged rows (Nz % 4 != 0 or unaligned buffers):
// four cells per thread moved as predicated 32-bit accesses (FDTDX_RAGGED, see common.cuh / yee_kernels.cuh).
#define FDTDX_RAGGED 1
#define FDTDX_BUILD_E 1
#include "yee_kernels.cuh"

template <int TIER, bool REV, int PM>
static void launch_E3(const StepParams& P, int t, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st) {
#define GO(S, A, M) yee_E_kernel<4, TIER, REV, S, A, M, PM><<<g, b, 0, st>>>(P, t)
  if constexpr (REV) {
    if (sig) { if (met) GO(true, false, true); else GO(true, false, false); }
    else { if (met) GO(false, false, true); else GO(false, false, false); }
  } else {
    if (ade) {
      if (sig) { if (met) GO(true, true, true); else GO(true, true, false); }
      else { if (met) GO(false, true, true); else GO(false, true, false); }
    } else {
      if (sig) { if (met) GO(true, false, true); else GO(true, false, false); }
      else { if (met) GO(false, false, true); else GO(false, false, false); }
    }
  }
#undef GO
}

template <int TIER, bool REV>
static void launch_E2(const StepParams& P, int t, int pm, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (pm == 0) launch_E3<TIER, REV, 0>(P, t, sig, ade, met, g, b, st);
  else if (pm == 1 || 1 == 1) launch_E3<TIER, REV, 1>(P, t, sig, ade, met, g, b, st);
  else launch_E3<TIER, REV, 1>(P, t, sig, ade, met, g, b, st);
}

void fdtdx_dispatch_E1(const StepParams& P, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (tier == 1) { if (rev) launch_E2<1, true>(P, t, pm, sig, ade, met, g, b, st); else launch_E2<1, false>(P, t, pm, sig, ade, met, g, b, st); }
  else { if (rev) launch_E2<3, true>(P, t, pm, sig, ade, met, g, b, st); else launch_E2<3, false>(P, t, pm, sig, ade, met, g, b, st); }
}
