// H half-step kernel instantiations + dispatch (see yee_kernels.cuh).
#define FDTDX_BUILD_H 1
#include "yee_kernels.cuh"

template <int V, int MUT>
static void launch_H2(const StepParams& P, int t, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st) {
#define GO(R, S, M) yee_H_kernel<V, MUT, R, S, M><<<g, b, 0, st>>>(P, t)
  if (rev) {
    if (sig) { if (met) GO(true, true, true); else GO(true, true, false); }
    else { if (met) GO(true, false, true); else GO(true, false, false); }
  } else {
    if (sig) { if (met) GO(false, true, true); else GO(false, true, false); }
    else { if (met) GO(false, false, true); else GO(false, false, false); }
  }
#undef GO
}

void fdtdx_dispatch_H(const StepParams& P, int t, bool v4, int mt, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (v4) {
    if (mt == 0) launch_H2<4, 0>(P, t, rev, sig, met, g, b, st);
    else if (mt == 1) launch_H2<4, 1>(P, t, rev, sig, met, g, b, st);
    else launch_H2<4, 3>(P, t, rev, sig, met, g, b, st);
  } else {
    if (mt == 0) launch_H2<1, 0>(P, t, rev, sig, met, g, b, st);
    else if (mt == 1) launch_H2<1, 1>(P, t, rev, sig, met, g, b, st);
    else launch_H2<1, 3>(P, t, rev, sig, met, g, b, st);
  }
}
