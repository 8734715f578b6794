// E half-step kernel instantiations + dispatch (see yee_kernels.cuh).
#define FDTDX_BUILD_E 1
#include "yee_kernels.cuh"

template <int V, int TIER, bool REV>
static void launch_E3(const StepParams& P, int t, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st) {
#define GO(S, A, M) yee_E_kernel<V, TIER, REV, S, A, M><<<g, b, 0, st>>>(P, t)
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

void fdtdx_dispatch_E(const StepParams& P, int t, bool v4, int tier, bool rev, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st) {
#define DISP(VV, TT)                                                         \
  do {                                                                       \
    if (rev) launch_E3<VV, TT, true>(P, t, sig, ade, met, g, b, st);         \
    else launch_E3<VV, TT, false>(P, t, sig, ade, met, g, b, st);            \
  } while (0)
  if (v4) { if (tier == 1) DISP(4, 1); else DISP(4, 3); }
  else { if (tier == 1) DISP(1, 1); else DISP(1, 3); }
#undef DISP
}
