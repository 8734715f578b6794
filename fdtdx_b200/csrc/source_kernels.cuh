// Source injection as its own O(surface) launch (tfsf.py:193-409, dipole.py:195-277).
//
// The half-step kernels inject sources in a cold pass of their own (src_pass_E/H) when the sources
// span at most a couple of x planes (x-normal TFSF / mode planes, dipoles): a CTA-uniform x test
// keeps every other plane free of injection code.  A y- or z-normal plane, however, crosses EVERY
// x plane, and the cold pass then serialises one lane per warp through the temporal profile on each
// plane (ncu: a 4.7 Mcell step went from 45 to 170 us).  Such sources run here instead: one thread
// per source cell, right after the forward half-step (before the reversed one), re-reading only the
// surface cells - the same O(surface) argument as the detectors.  Arithmetic and per-cell order of
// sources are those of inject_E / inject_H, so results are bit-identical to the in-kernel pass.
#pragma once
#include "yee_kernels.cuh"

// grid: (ceil(max box cells / 256), n_src).  A cell covered by several source boxes is handled once,
// by the thread of the first covering source, which applies all sources in order (like the reference's
// loop over objects.sources).
template <int TIER_E, int TIER_M>
__global__ void src_apply_kernel(const StepParams P, const int t, const int is_E, const int reverse) {
  const int s = blockIdx.y;
  // the source descriptors live in global memory; stage them once per block (they are read many times
  // along a dependent chain: box, switch tables, profile parameters, array pointers)
  __shared__ SrcDev s_src[FDTDX_MAX_SRC];
  pdl_trigger();
  for (int q = threadIdx.x; q < P.n_src * (int)(sizeof(SrcDev) / 4); q += blockDim.x)
    reinterpret_cast<int*>(s_src)[q] = reinterpret_cast<const int*>(P.src)[q];
  __syncthreads();
  pdl_wait();
  const int lo0 = max(P.src_lo[s][0], P.x_begin), hi0 = min(P.src_hi[s][0], P.x_end);
  const int d0 = hi0 - lo0, d1 = P.src_hi[s][1] - P.src_lo[s][1], d2 = P.src_hi[s][2] - P.src_lo[s][2];
  if (d0 <= 0 || d1 <= 0 || d2 <= 0) return;
  const long long n = (long long)d0 * d1 * d2;
  const long long plane = (long long)P.ny * P.nz, N = plane * P.nx;
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < n; f += (long long)gridDim.x * blockDim.x) {
    const int k = P.src_lo[s][2] + (int)(f % d2);
    const int j = P.src_lo[s][1] + (int)((f / d2) % d1);
    const int i = lo0 + (int)(f / ((long long)d2 * d1));
    bool first = true;
    for (int q = 0; q < s; ++q)
      if (i >= P.src_lo[q][0] && i < P.src_hi[q][0] && j >= P.src_lo[q][1] && j < P.src_hi[q][1] && k >= P.src_lo[q][2] && k < P.src_hi[q][2]) first = false;
    if (!first) continue;
    const long long cell = (long long)i * plane + (long long)j * P.nz + k;
    float* F = is_E ? (P.E_out ? P.E_out : P.E) : (P.H_out ? P.H_out : P.H);
    float f0 = F[cell], f1 = F[N + cell], f2 = F[2 * N + cell];
    float m0, m1, m2;
    if (is_E) {
      m0 = P.eps[cell];
      m1 = (TIER_E == 3) ? P.eps[P.eps_cs + cell] : m0;
      m2 = (TIER_E == 3) ? P.eps[2 * P.eps_cs + cell] : m0;
      inject_E(s_src, P.n_src, P.dt, t, reverse != 0, i, j, k, m0, m1, m2, &f0, &f1, &f2);
    } else {
      m0 = m1 = m2 = P.inv_mu_scalar;
      if (TIER_M >= 1) {
        m0 = P.mu[cell];
        m1 = (TIER_M == 3) ? P.mu[P.mu_cs + cell] : m0;
        m2 = (TIER_M == 3) ? P.mu[2 * P.mu_cs + cell] : m0;
      }
      inject_H(s_src, P.n_src, P.dt, t, reverse != 0, i, j, k, m0, m1, m2, &f0, &f1, &f2);
    }
    if (!reverse) {
      // the reference masks PEC / PMC cells after the injection (update.py order)
      for (int w = 0; w < P.n_walls; ++w) {
        const WallDev& W = P.wallp[w];
        if (W.kind == (is_E ? 0 : 1) && in_box(W.lo, W.hi, i, j, k)) {
          if (W.axis != 0) f0 = 0.0f;
          if (W.axis != 1) f1 = 0.0f;
          if (W.axis != 2) f2 = 0.0f;
        }
      }
    }
    F[cell] = f0; F[N + cell] = f1; F[2 * N + cell] = f2;
  }
}
