// H half-step kernel instantiations + dispatch for V = 4 (see yee_kernels.cuh).
#define FDTDX_BUILD_H 1
#include "yee_kernels.cuh"

template <int MUT, int PM>
static void launch_H3(const StepParams& P, int t, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st) {
#define GO(R, S, M) yee_H_kernel<4, MUT, R, S, M, PM><<<g, b, 0, st>>>(P, t)
  if (rev) {
    if (sig) { if (met) GO(true, true, true); else GO(true, true, false); }
    else { if (met) GO(true, false, true); else GO(true, false, false); }
  } else {
    if (sig) { if (met) GO(false, true, true); else GO(false, true, false); }
    else { if (met) GO(false, false, true); else GO(false, false, false); }
  }
#undef GO
}

template <int MUT>
static void launch_H2(const StepParams& P, int t, int pm, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (pm == 0) launch_H3<MUT, 0>(P, t, rev, sig, met, g, b, st);
  else if (pm == 1 || 4 == 1) launch_H3<MUT, 1>(P, t, rev, sig, met, g, b, st);
  else launch_H3<MUT, 2>(P, t, rev, sig, met, g, b, st);
}

void fdtdx_dispatch_H4(const StepParams& P, int t, int mt, int pm, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (mt == 0) launch_H2<0>(P, t, pm, rev, sig, met, g, b, st);
  else if (mt == 1) launch_H2<1>(P, t, pm, rev, sig, met, g, b, st);
  else launch_H2<3>(P, t, pm, rev, sig, met, g, b, st);
}

// curl-only launches (phase 1 of the full-tensor tier): K = curl +- CPML written to P.H
template <bool REV_, bool MET_, int PM_>
static void konly_go(const StepParams& P, int t, dim3 g, dim3 b, cudaStream_t st) {
  yee_H_kernel<4, 0, REV_, false, MET_, PM_, true><<<g, b, 0, st>>>(P, t);
}
template <bool REV_, bool MET_>
static void konly_pm(const StepParams& P, int t, int pm, dim3 g, dim3 b, cudaStream_t st) {
  if (pm == 0) konly_go<REV_, MET_, 0>(P, t, g, b, st);
  else if (pm == 1) konly_go<REV_, MET_, 1>(P, t, g, b, st);
  else konly_go<REV_, MET_, 2>(P, t, g, b, st);
}
void fdtdx_dispatch_H4_konly(const StepParams& P, int t, int pm, bool rev, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (rev) { if (met) konly_pm<true, true>(P, t, pm, g, b, st); else konly_pm<true, false>(P, t, pm, g, b, st); }
  else { if (met) konly_pm<false, true>(P, t, pm, g, b, st); else konly_pm<false, false>(P, t, pm, g, b, st); }
}
