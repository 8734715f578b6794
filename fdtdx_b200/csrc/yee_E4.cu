// E half-step kernel instantiations + dispatch for V = 4 (see yee_kernels.cuh).
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
  else if (pm == 1 || 4 == 1) launch_E3<TIER, REV, 1>(P, t, sig, ade, met, g, b, st);
  else launch_E3<TIER, REV, 2>(P, t, sig, ade, met, g, b, st);
}

void fdtdx_dispatch_E4(const StepParams& P, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (tier == 1) { if (rev) launch_E2<1, true>(P, t, pm, sig, ade, met, g, b, st); else launch_E2<1, false>(P, t, pm, sig, ade, met, g, b, st); }
  else { if (rev) launch_E2<3, true>(P, t, pm, sig, ade, met, g, b, st); else launch_E2<3, false>(P, t, pm, sig, ade, met, g, b, st); }
}

// curl-only launches (phase 1 of the full-tensor tier): K = curl +- CPML written to P.E
template <bool REV_, bool MET_, int PM_>
static void konly_go(const StepParams& P, int t, dim3 g, dim3 b, cudaStream_t st) {
  yee_E_kernel<4, 1, REV_, false, false, MET_, PM_, true><<<g, b, 0, st>>>(P, t);
}
template <bool REV_, bool MET_>
static void konly_pm(const StepParams& P, int t, int pm, dim3 g, dim3 b, cudaStream_t st) {
  if (pm == 0) konly_go<REV_, MET_, 0>(P, t, g, b, st);
  else if (pm == 1) konly_go<REV_, MET_, 1>(P, t, g, b, st);
  else konly_go<REV_, MET_, 2>(P, t, g, b, st);
}
void fdtdx_dispatch_E4_konly(const StepParams& P, int t, int pm, bool rev, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (rev) { if (met) konly_pm<true, true>(P, t, pm, g, b, st); else konly_pm<true, false>(P, t, pm, g, b, st); }
  else { if (met) konly_pm<false, true>(P, t, pm, g, b, st); else konly_pm<false, false>(P, t, pm, g, b, st); }
}
