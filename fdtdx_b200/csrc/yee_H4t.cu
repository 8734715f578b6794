// TMA-staged H half-step: instantiations + dispatch (see yee_tma.cuh).
#define FDTDX_BUILD_H 1
#include "yee_tma.cuh"
#include "tma_cfg.h"
#ifndef FDTDX_TZ_SEL
#define FDTDX_TZ_SEL 128  // tile width of this translation unit (yee_H4t64.cu re-includes this file with 64)
#endif
#include <cstdlib>

static bool fdtdx_tma_pdl_enabled() {
  const char* e = getenv("FDTDX_B200_PDL");
  return !(e && e[0] == '0');
}

template <int MUT, bool REV, bool SIG, bool MET, int PM>
static cudaError_t go_H(const StepParams& P, const TmaSet& M, int t, dim3 g, cudaStream_t st) {
  // three material components make a stage 39 KB: a 3-deep ring would leave one CTA per SM, so those
  // variants run a 2-deep ring (78 KB, two CTAs per SM)
  constexpr int R = FDTDX_TMA_R, S = (MUT == 3) ? 2 : FDTDX_TMA_S;
  constexpr int TZ = FDTDX_TZ_SEL;
  // flat tiles (TZ == 0): the geometry, and with it the shared-memory size, follows the row length of the grid
  const int smem = TZ > 0 ? tma_smem_bytes<R, (TZ > 0 ? TZ : 128), MUT, S>() : tma_smem_bytes_flat<R>(P.flat_lz, MUT, S);
  auto k = yee_H_tma<MUT, REV, SIG, MET, PM, R, S, TZ>;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr_smem = smem;
  }
  // programmatic dependent launch: this grid may begin while the previous kernel of the stream drains
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g;
  cfg.blockDim = dim3(32, R);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = fdtdx_tma_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, P, M, t);
}

template <int MUT, int PM>
static cudaError_t launch_H3(const StepParams& P, const TmaSet& M, int t, bool rev, bool sig, bool met, dim3 g, cudaStream_t st) {
#define GO(R_, S_, M_) return go_H<MUT, R_, S_, M_, PM>(P, M, t, g, st)
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
static cudaError_t launch_H2(const StepParams& P, const TmaSet& M, int t, int pm, bool rev, bool sig, bool met, dim3 g, cudaStream_t st) {
  if (pm == 0) return launch_H3<MUT, 0>(P, M, t, rev, sig, met, g, st);
  if (pm == 1) return launch_H3<MUT, 1>(P, M, t, rev, sig, met, g, st);
  return launch_H3<MUT, 2>(P, M, t, rev, sig, met, g, st);
}

cudaError_t fdtdx_dispatch_H4_tma(const StepParams& P, const TmaSet& M, int t, int mt, int pm, bool rev, bool sig, bool met, dim3 g, cudaStream_t st) {
  if (mt == 0) return launch_H2<0>(P, M, t, pm, rev, sig, met, g, st);
  if (mt == 1) return launch_H2<1>(P, M, t, pm, rev, sig, met, g, st);
  return launch_H2<3>(P, M, t, pm, rev, sig, met, g, st);
}

#if FDTDX_TZ_SEL == 128
// curl-only mode (phase 1 of the full-tensor tier): K = curl + CPML correction, written to P.H
template <bool REV, bool MET, int PM>
static cudaError_t go_H_konly(const StepParams& P, const TmaSet& M, int t, dim3 g, cudaStream_t st) {
  constexpr int R = FDTDX_TMA_R, S = FDTDX_TMA_S, TZ = FDTDX_TZ_SEL;
  constexpr int smem = tma_smem_bytes<R, TZ, 0, S>();
  auto k = yee_H_tma<0, REV, false, MET, PM, R, S, TZ, true>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g;
  cfg.blockDim = dim3(32, R);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 0;  // the neighbours of this launch are plain launches
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, P, M, t);
}
template <bool REV, bool MET>
static cudaError_t konly_pm_H(const StepParams& P, const TmaSet& M, int t, int pm, dim3 g, cudaStream_t st) {
  if (pm == 0) return go_H_konly<REV, MET, 0>(P, M, t, g, st);
  if (pm == 1) return go_H_konly<REV, MET, 1>(P, M, t, g, st);
  return go_H_konly<REV, MET, 2>(P, M, t, g, st);
}
cudaError_t fdtdx_dispatch_H4_tma_konly(const StepParams& P, const TmaSet& M, int t, int pm, bool rev, bool met, dim3 g, cudaStream_t st) {
  if (rev) return met ? konly_pm_H<true, true>(P, M, t, pm, g, st) : konly_pm_H<true, false>(P, M, t, pm, g, st);
  return met ? konly_pm_H<false, true>(P, M, t, pm, g, st) : konly_pm_H<false, false>(P, M, t, pm, g, st);
}
#endif
