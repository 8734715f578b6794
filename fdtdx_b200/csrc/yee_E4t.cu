// TMA-staged E half-step: instantiations + dispatch (see yee_tma.cuh).
#define FDTDX_BUILD_E 1
#include "yee_tma.cuh"
#include "tma_cfg.h"
#ifndef FDTDX_TZ_SEL
#define FDTDX_TZ_SEL 128  // tile width of this translation unit (yee_E4t64.cu re-includes this file with 64)
#endif
#include <cstdlib>

static bool fdtdx_tma_pdl_enabled() {
  const char* e = getenv("FDTDX_B200_PDL");
  return !(e && e[0] == '0');
}

template <int TIER, bool REV, bool SIG, bool ADE, bool MET, int PM>
static cudaError_t go_E(const StepParams& P, const TmaSet& M, int t, dim3 g, cudaStream_t st) {
  // three material components make a stage 39 KB: a 3-deep ring would leave one CTA per SM, so those
  // variants run a 2-deep ring (78 KB, two CTAs per SM)
  constexpr int R = FDTDX_TMA_R, S = (TIER == 3) ? 2 : FDTDX_TMA_S;
  constexpr int TZ = FDTDX_TZ_SEL;
  // flat tiles (TZ == 0): the geometry, and with it the shared-memory size, follows the row length of the grid
  const int smem = TZ > 0 ? tma_smem_bytes<R, (TZ > 0 ? TZ : 128), TIER, S>() : tma_smem_bytes_flat<R>(P.flat_lz, TIER, S);
  auto k = yee_E_tma<TIER, REV, SIG, ADE, MET, PM, R, S, TZ>;
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

template <int TIER, bool REV, int PM>
static cudaError_t launch_E3(const StepParams& P, const TmaSet& M, int t, bool sig, bool ade, bool met, dim3 g, cudaStream_t st) {
#define GO(S_, A_, M_) return go_E<TIER, REV, S_, A_, M_, PM>(P, M, t, g, st)
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
static cudaError_t launch_E2(const StepParams& P, const TmaSet& M, int t, int pm, bool sig, bool ade, bool met, dim3 g, cudaStream_t st) {
  if (pm == 0) return launch_E3<TIER, REV, 0>(P, M, t, sig, ade, met, g, st);
  if (pm == 1) return launch_E3<TIER, REV, 1>(P, M, t, sig, ade, met, g, st);
  return launch_E3<TIER, REV, 2>(P, M, t, sig, ade, met, g, st);
}

cudaError_t fdtdx_dispatch_E4_tma(const StepParams& P, const TmaSet& M, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g,
                                  cudaStream_t st) {
  if (tier == 1) return rev ? launch_E2<1, true>(P, M, t, pm, sig, ade, met, g, st) : launch_E2<1, false>(P, M, t, pm, sig, ade, met, g, st);
  return rev ? launch_E2<3, true>(P, M, t, pm, sig, ade, met, g, st) : launch_E2<3, false>(P, M, t, pm, sig, ade, met, g, st);
}

#if FDTDX_TZ_SEL == 128
// curl-only mode (phase 1 of the full-tensor tier): K = curl + CPML correction, written to P.E
template <bool REV, bool MET, int PM>
static cudaError_t go_E_konly(const StepParams& P, const TmaSet& M, int t, dim3 g, cudaStream_t st) {
  constexpr int R = FDTDX_TMA_R, S = FDTDX_TMA_S, TZ = FDTDX_TZ_SEL;
  constexpr int smem = tma_smem_bytes<R, TZ, 1, S>();
  auto k = yee_E_tma<1, REV, false, false, MET, PM, R, S, TZ, true>;
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
static cudaError_t konly_pm_E(const StepParams& P, const TmaSet& M, int t, int pm, dim3 g, cudaStream_t st) {
  if (pm == 0) return go_E_konly<REV, MET, 0>(P, M, t, g, st);
  if (pm == 1) return go_E_konly<REV, MET, 1>(P, M, t, g, st);
  return go_E_konly<REV, MET, 2>(P, M, t, g, st);
}
cudaError_t fdtdx_dispatch_E4_tma_konly(const StepParams& P, const TmaSet& M, int t, int pm, bool rev, bool met, dim3 g, cudaStream_t st) {
  if (rev) return met ? konly_pm_E<true, true>(P, M, t, pm, g, st) : konly_pm_E<true, false>(P, M, t, pm, g, st);
  return met ? konly_pm_E<false, true>(P, M, t, pm, g, st) : konly_pm_E<false, false>(P, M, t, pm, g, st);
}
#endif
