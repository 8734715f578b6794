// H half-step kernel instantiations + dispatch for ragged rows (Nz % 4 != 0 or unaligned buffers):
// four cells per thread moved as predicated 32-bit accesses (FDTDX_RAGGED, see common.cuh / yee_kernels.cuh).
#define FDTDX_RAGGED 1
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
  else if (pm == 1 || 1 == 1) launch_H3<MUT, 1>(P, t, rev, sig, met, g, b, st);
  else launch_H3<MUT, 1>(P, t, rev, sig, met, g, b, st);
}

void fdtdx_dispatch_H1(const StepParams& P, int t, int mt, int pm, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st) {
  if (mt == 0) launch_H2<0>(P, t, pm, rev, sig, met, g, b, st);
  else if (mt == 1) launch_H2<1>(P, t, pm, rev, sig, met, g, b, st);
  else launch_H2<3>(P, t, pm, rev, sig, met, g, b, st);
}
