// Yee E / H half-step kernels for sm_100a: isotropic and diagonal materials.
//
// One kernel per half-step fuses what the reference spreads over pad -> diff -> scale -> CPML
// scatter -> material update -> ADE -> source injection -> PEC/PMC masking:
//   yee_E_kernel : pad_fields_for_boundaries(H) + curl_H + step_cpml + update_E / update_E_reverse
//                  (fdtd/update.py:92-136, 256-354, 494-523, 526-609; core/physics/curl.py:314-397)
//   yee_H_kernel : pad_fields_for_boundaries(E) + curl_E + step_cpml + update_H / update_H_reverse
//                  (fdtd/update.py:689-750, 824-853, 856-930; core/physics/curl.py:227-311)
//
// Mapping.  Arrays are the reference's (3,Nx,Ny,Nz) float32, z fastest.  A warp owns 32*V
// consecutive z cells of one y row (V = 4: one 128-bit load per component per thread); a CTA owns
// ROWS y rows and marches along x through a chunk of planes.  The x-neighbour plane lives in a
// register queue (each field value is loaded from HBM once per chunk), the z-neighbour comes from
// the adjacent lane by warp shuffle, the y-neighbour row is re-read through L1 (it is the row the
// neighbouring warp of the same CTA just loaded).  Zero / wrap / neighbour-rank halos are resolved
// by index, so no padded copy is ever materialised.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// temporal profiles (objects/sources/profile.py:263-273, 322-345, 412-439; core/window.py:16-30)
// float32, same operation order as the reference; no FMA contraction (compiled with -fmad=false).
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ float src_profile_ph(const SrcDev& S, float time, const float ph1) {
  if (S.profile_kind == 0) {  // SingleFrequencyProfile
    float phase = ((S.p[4] * time) / S.p[0] + ph1) + S.p[2];  // 2*pi*time/period + phase_shift + self.phase_shift
    float raw = cosf(phase);                                     // Re(exp(-i*phase))
    float f = time / S.p[3];
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    return f * raw;
  } else if (S.profile_kind == 1) {  // GaussianPulseProfile
    float d = time - S.p[3];
    float env = expf(-(d * d) / S.p[5]);
    float phase = (S.p[0] * time + ph1) + S.p[2];
    return env * cosf(phase);
  } else {  // CustomTimeSignalProfile
    float idx = (time - S.p[0]) / S.p[1];
    float fl = floorf(idx);
    int i0 = (int)fl;
    float frac = idx - fl;
    int n = S.signal_len;
    bool valid = (i0 >= 0) && (i0 < n);
    int a = min(max(i0, 0), n - 1);
    int b = min(max(a + 1, 0), n - 1);
    float y0 = S.signal[a], y1 = S.signal[b];
    float y = (S.p[3] != 0.0f) ? (frac < 0.5f ? y0 : y1) : ((1.0f - frac) * y0 + frac * y1);
    return valid ? y : S.p[2];
  }
}

__device__ __forceinline__ float src_profile(const SrcDev& S, float time) { return src_profile_ph(S, time, S.p[1]); }

__device__ __forceinline__ bool src_time(const SrcDev& S, int t, float half, float* tf) {
  if (S.on != nullptr) {
    if (!S.on[t]) return false;
    *tf = S.t_adj[t] + half;
  } else {
    *tf = (float)t + half;
  }
  return true;
}

__device__ __forceinline__ bool in_box(const int* lo, const int* hi, int i, int j, int k) {
  return i >= lo[0] && i < hi[0] && j >= lo[1] && j < hi[1] && k >= lo[2] && k < hi[2];
}

// E-side injection at one cell (tfsf.py:193-308 diagonal branch; dipole.py:195-232).
static __device__ __noinline__ void inject_E(const SrcDev* __restrict__ srcs, int n_src, float dt, int t, bool reverse, int i, int j, int k,
                                       float ie0, float ie1, float ie2, float* e0, float* e1, float* e2) {
  float ie[3] = {ie0, ie1, ie2};
  float E[3] = {*e0, *e1, *e2};
  for (int s = 0; s < n_src; ++s) {
    const SrcDev& S = srcs[s];
    if (!in_box(S.lo, S.hi, i, j, k)) continue;
    float tf;
    if (!src_time(S, t, 0.0f, &tf)) continue;
    if (S.kind == 0) {
      int fy = S.hi[1] - S.lo[1], fz = S.hi[2] - S.lo[2];
      long long fn = (long long)(S.hi[0] - S.lo[0]) * fy * fz;
      long long f = ((long long)(i - S.lo[0]) * fy + (j - S.lo[1])) * fz + (k - S.lo[2]);
      int a = (S.normal_axis + 1) % 3, b = (S.normal_axis + 2) % 3;
      float sign = reverse ? -S.sign : S.sign;
      const float inc_b = S.Hinc[b * fn + f], inc_a = S.Hinc[a * fn + f];  // loaded ahead of the profile calls
      float amp_a, amp_b;
      if (S.hfilter == nullptr) {
        amp_a = src_profile(S, (tf + S.toffH[a * fn + f]) * dt) * S.static_amp;
        amp_b = src_profile(S, (tf + S.toffH[b * fn + f]) * dt) * S.static_amp;
      } else {
        // jnp.interp(t + toff, arange(T), filter, left=0, right=0)   (tfsf.py:259-264)
        float amps[2];
        int ax2[2] = {a, b};
        for (int q = 0; q < 2; ++q) {
          float x = tf + S.toffH[ax2[q] * fn + f];
          float v = 0.0f;
          if (x >= 0.0f && x <= (float)(S.hfilter_len - 1)) {
            int i0 = min((int)floorf(x), S.hfilter_len - 2);
            i0 = max(i0, 0);
            float f0 = S.hfilter[i0], f1 = S.hfilter[i0 + 1];
            v = f0 + (x - (float)i0) * (f1 - f0);
          }
          amps[q] = v * S.static_amp;
        }
        amp_a = amps[0];
        amp_b = amps[1];
      }
      float Hb = inc_b * amp_b;
      float Ha = inc_a * amp_a;
      if (S.HincI != nullptr && S.hfilter == nullptr) {  // Re * amp + Im * amp_quadrature
        const float aq_a = src_profile_ph(S, (tf + S.toffH[a * fn + f]) * dt, S.pq) * S.static_amp;
        const float aq_b = src_profile_ph(S, (tf + S.toffH[b * fn + f]) * dt, S.pq) * S.static_amp;
        Hb = Hb + S.HincI[b * fn + f] * aq_b;
        Ha = Ha + S.HincI[a * fn + f] * aq_a;
      }
      Hb = (Hb * S.cE) * ie[a];
      Ha = (Ha * S.cE) * ie[b];
      E[a] = E[a] + sign * Hb;
      E[b] = E[b] + (-sign) * Ha;
    } else if (S.electric) {
      float amp = src_profile(S, tf * dt);
      float sg = reverse ? 1.0f : -1.0f;
      float scale = S.dip_scale * amp;
      E[S.pol] = E[S.pol] + sg * (scale * ie[S.pol]);
    }
  }
  *e0 = E[0]; *e1 = E[1]; *e2 = E[2];
}

// H-side injection at one cell (tfsf.py:311-409 diagonal branch; dipole.py:236-277).
static __device__ __noinline__ void inject_H(const SrcDev* __restrict__ srcs, int n_src, float dt, int t, bool reverse, int i, int j, int k,
                                       float im0, float im1, float im2, float* h0, float* h1, float* h2) {
  float im[3] = {im0, im1, im2};
  float H[3] = {*h0, *h1, *h2};
  for (int s = 0; s < n_src; ++s) {
    const SrcDev& S = srcs[s];
    if (!in_box(S.lo, S.hi, i, j, k)) continue;
    float tf;
    if (!src_time(S, t, 0.5f, &tf)) continue;
    if (S.kind == 0) {
      int fy = S.hi[1] - S.lo[1], fz = S.hi[2] - S.lo[2];
      long long fn = (long long)(S.hi[0] - S.lo[0]) * fy * fz;
      long long f = ((long long)(i - S.lo[0]) * fy + (j - S.lo[1])) * fz + (k - S.lo[2]);
      int a = (S.normal_axis + 1) % 3, b = (S.normal_axis + 2) % 3;
      float sign = reverse ? -S.sign : S.sign;
      const float inc_a = S.Einc[a * fn + f], inc_b = S.Einc[b * fn + f];
      const float to_a = S.toffE[a * fn + f], to_b = S.toffE[b * fn + f];
      float amp_a = src_profile(S, (tf + to_a) * dt) * S.static_amp;
      float amp_b = src_profile(S, (tf + to_b) * dt) * S.static_amp;
      float Ea = inc_a * amp_a;
      float Eb = inc_b * amp_b;
      if (S.EincI != nullptr) {
        const float aq_a = src_profile_ph(S, (tf + to_a) * dt, S.pq) * S.static_amp;
        const float aq_b = src_profile_ph(S, (tf + to_b) * dt, S.pq) * S.static_amp;
        Ea = Ea + S.EincI[a * fn + f] * aq_a;
        Eb = Eb + S.EincI[b * fn + f] * aq_b;
      }
      Ea = (Ea * S.cH) * im[b];
      Eb = (Eb * S.cH) * im[a];
      H[b] = H[b] + sign * Ea;
      H[a] = H[a] + (-sign) * Eb;
    } else if (!S.electric) {
      float amp = src_profile(S, tf * dt);
      float sg = reverse ? 1.0f : -1.0f;
      float scale = S.dip_scale * amp;
      H[S.pol] = H[S.pol] + sg * (scale * im[S.pol]);
    }
  }
  *h0 = H[0]; *h1 = H[1]; *h2 = H[2];
}

// ------------------------------------------------------------------------------------------------
// Peer-memory halo ordering (StepParams::peer_*).  One rule per half-step covers the read-after-write
// and the write-after-read hazard on the plane shared with a neighbour.
// ------------------------------------------------------------------------------------------------
// Called by every thread of a boundary-chunk CTA before its first access to the shared plane: one
// thread polls (acquire, system scope), the CTA barrier publishes the result.  Bounded: a neighbour
// that never arrives (a rank that died or issued fewer half-steps) raises peer_err instead of hanging.
__device__ __forceinline__ void peer_wait_cta(const StepParams& P) {
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    unsigned ns = 32;
    const long long t0 = clock64();
    for (;;) {
      int v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(P.peer_wait) : "memory");
      if (v - P.peer_wait_target >= 0) break;
      if (clock64() - t0 > (1LL << 36)) {  // ~35 s at 1.9 GHz
        atomicExch(P.peer_err, 1);
        break;
      }
      __nanosleep(ns);
      if (ns < 1024) ns *= 2;
    }
  }
  __syncthreads();
}
// Called by the active lanes of a boundary-chunk warp after its last store.
__device__ __forceinline__ void peer_signal_warp(const StepParams& P, const unsigned mask) {
  __threadfence_system();  // this lane's stores are visible system-wide before the arrival
  __syncwarp(mask);
  if ((threadIdx.x & 31) == __ffs(mask) - 1) {
    const int old = atomicAdd(P.peer_ctr, 1);
    if (old == P.peer_total - 1) {
      atomicExch(P.peer_ctr, 0);
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(P.peer_signal), "r"(P.peer_signal_value) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Hot-loop helpers.  Everything the marching loop touches per plane lives in registers or in kernel
// parameter (constant) space; source injection and wall masking are kept out of line so that the
// common cell costs no local-memory traffic and no descriptor loads.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
#ifndef FDTDX_PF_DIST
#define FDTDX_PF_DIST 2
#endif
#ifndef FDTDX_MIN_CTAS
#define FDTDX_MIN_CTAS 2
#endif

// Does any source box intersect cells (i, j, k0..k0+V-1)?  Boxes come from parameter space.
__device__ __forceinline__ bool src_hits(const StepParams& P, int i, int j, int k0, int V) {
  if (i < P.src_x0 || i >= P.src_x1) return false;  // CTA-uniform: almost every plane stops here
  for (int s = 0; s < P.n_src; ++s) {
    if (i >= P.src_lo[s][0] && i < P.src_hi[s][0] && j >= P.src_lo[s][1] && j < P.src_hi[s][1] && k0 < P.src_hi[s][2] &&
        k0 + (V - 1) * FDTDX_ES + 1 > P.src_lo[s][2])
      return true;
  }
  return false;
}

// Cold path, outside the marching loop: inject the sources into this thread's cells of planes
// [ic0, ic1) in global memory.  Forward: after the loop (the cell already holds the updated field;
// the wall mask is re-applied because the reference masks after the injection).  Reverse: before the
// loop (update_E_reverse undoes the injection first).  Same arithmetic as a fused epilogue; a thread
// only touches cells it owns, so no other thread's data is involved.
template <int V, int TIER>
static __device__ __noinline__ void src_pass_E(const StepParams& P, int t, bool reverse, int ic0, int ic1, int j, int k0) {
  const long long plane = (long long)P.ny * P.nz;
  const long long N = plane * P.nx;
  for (int i = max(ic0, P.src_x0); i < min(ic1, P.src_x1); ++i) {
    if (!src_hits(P, i, j, k0, V)) continue;
    const long long cell0 = (long long)i * plane + (long long)j * P.nz + k0;
    for (int e = 0; e < V && k0 + e * FDTDX_ES < P.nz; ++e) {
      const long long cell = cell0 + e * FDTDX_ES;
      float* const Ed = P.E_out ? P.E_out : P.E;
      float e0 = Ed[cell], e1 = Ed[N + cell], e2 = Ed[2 * N + cell];
      const float i0 = P.eps[cell];
      const float i1 = (TIER == 3) ? P.eps[P.eps_cs + cell] : i0;
      const float i2 = (TIER == 3) ? P.eps[2 * P.eps_cs + cell] : i0;
      inject_E(P.src, P.n_src, P.dt, t, reverse, i, j, k0 + e * FDTDX_ES, i0, i1, i2, &e0, &e1, &e2);
      if (!reverse) {
        for (int w = 0; w < P.n_walls; ++w) {
          const WallDev& W = P.wallp[w];
          if (W.kind == 0 && in_box(W.lo, W.hi, i, j, k0 + e * FDTDX_ES)) {
            if (W.axis != 0) e0 = 0.0f;
            if (W.axis != 1) e1 = 0.0f;
            if (W.axis != 2) e2 = 0.0f;
          }
        }
      }
      Ed[cell] = e0; Ed[N + cell] = e1; Ed[2 * N + cell] = e2;
    }
  }
}
template <int V, int MUT>
static __device__ __noinline__ void src_pass_H(const StepParams& P, int t, bool reverse, int ic0, int ic1, int j, int k0) {
  const long long plane = (long long)P.ny * P.nz;
  const long long N = plane * P.nx;
  for (int i = max(ic0, P.src_x0); i < min(ic1, P.src_x1); ++i) {
    if (!src_hits(P, i, j, k0, V)) continue;
    const long long cell0 = (long long)i * plane + (long long)j * P.nz + k0;
    for (int e = 0; e < V && k0 + e * FDTDX_ES < P.nz; ++e) {
      const long long cell = cell0 + e * FDTDX_ES;
      float* const Hd = P.H_out ? P.H_out : P.H;
      float h0 = Hd[cell], h1 = Hd[N + cell], h2 = Hd[2 * N + cell];
      float m0 = P.inv_mu_scalar, m1 = m0, m2 = m0;
      if (MUT >= 1) {
        m0 = P.mu[cell];
        m1 = (MUT == 3) ? P.mu[P.mu_cs + cell] : m0;
        m2 = (MUT == 3) ? P.mu[2 * P.mu_cs + cell] : m0;
      }
      inject_H(P.src, P.n_src, P.dt, t, reverse, i, j, k0 + e * FDTDX_ES, m0, m1, m2, &h0, &h1, &h2);
      if (!reverse) {
        for (int w = 0; w < P.n_walls; ++w) {
          const WallDev& W = P.wallp[w];
          if (W.kind == 1 && in_box(W.lo, W.hi, i, j, k0 + e * FDTDX_ES)) {
            if (W.axis != 0) h0 = 0.0f;
            if (W.axis != 1) h1 = 0.0f;
            if (W.axis != 2) h2 = 0.0f;
          }
        }
      }
      Hd[cell] = h0; Hd[N + cell] = h1; Hd[2 * N + cell] = h2;
    }
  }
}

// PEC (kind 0) / PMC (kind 1) tangential zeroing of V freshly computed cells (pec.py:70-77, pmc.py:63-76).
template <int V>
__device__ __forceinline__ void wall_mask(const StepParams& P, int kind, int i, int j, int k0, Vec<V>& o0, Vec<V>& o1, Vec<V>& o2) {
  for (int w = 0; w < P.n_walls; ++w) {
    const WallDev& W = P.wallp[w];
    if (W.kind != kind || i < W.lo[0] || i >= W.hi[0] || j < W.lo[1] || j >= W.hi[1]) continue;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      if (k0 + e * FDTDX_ES >= W.lo[2] && k0 + e * FDTDX_ES < W.hi[2]) {
        if (W.axis != 0) o0.v[e] = 0.0f;
        if (W.axis != 1) o1.v[e] = 0.0f;
        if (W.axis != 2) o2.v[e] = 0.0f;
      }
    }
  }
}

// CPML for one axis on V cells (perfectly_matched_layer.py:138-190; curl.py:284-308, 371-394).
// d1 = d_a F_j, d2 = d_a F_i; psi' = b psi + a d (UPD), correction = (1/kappa - 1) d + psi' (K1: kappa == 1,
// correction = psi'); Km -= corr_1, Kp += corr_2.
template <int V, bool UPD, bool K1>
__device__ __forceinline__ void cpml_axis(const float a, const float b, const float km1, const Vec<V>& d1, const Vec<V>& d2,
                                          Vec<V>& p1, Vec<V>& p2, Vec<V>& Km, Vec<V>& Kp) {
#pragma unroll
  for (int e = 0; e < V; ++e) {
    float q1 = p1.v[e], q2 = p2.v[e];
    if (UPD) {
      q1 = b * q1 + a * d1.v[e];
      q2 = b * q2 + a * d2.v[e];
      p1.v[e] = q1;
      p2.v[e] = q2;
    }
    float c1 = q1, c2 = q2;
    if (!K1) {
      c1 = km1 * d1.v[e] + q1;
      c2 = km1 * d2.v[e] + q2;
    }
    Km.v[e] = Km.v[e] - c1;
    Kp.v[e] = Kp.v[e] + c2;
  }
}
// vector coefficients (z axis: one coefficient per cell), elements [E0, E1)
template <int V, int E0, int E1, bool UPD, bool K1>
__device__ __forceinline__ void cpml_axis_v(const Vec<V>& a, const Vec<V>& b, const Vec<V>& km1, const Vec<V>& d1, const Vec<V>& d2,
                                            Vec<V>& p1, Vec<V>& p2, Vec<V>& Km, Vec<V>& Kp) {
#pragma unroll
  for (int e = E0; e < E1; ++e) {
    float q1 = p1.v[e], q2 = p2.v[e];
    if (UPD) {
      q1 = b.v[e] * q1 + a.v[e] * d1.v[e];
      q2 = b.v[e] * q2 + a.v[e] * d2.v[e];
      p1.v[e] = q1;
      p2.v[e] = q2;
    }
    float c1 = q1, c2 = q2;
    if (!K1) {
      c1 = km1.v[e] * d1.v[e] + q1;
      c2 = km1.v[e] * d2.v[e] + q2;
    }
    Km.v[e] = Km.v[e] - c1;
    Kp.v[e] = Kp.v[e] + c2;
  }
}

// Per-thread, loop-invariant description of the y / z CPML slabs this thread's cells belong to.
template <int V>
struct PmlLane {
  bool in_y, any_z, zvec, zh0, zh1;
  int zm;  // aligned kernels, PM == 1: bit e set = cell k0+e lies in this lane's z slab
  long long ystride, yoff;  // psi index of plane i: i * ystride + yoff
  long long zstride, zoff;
  float ay, by, ky;
  Vec<V> az, bz, kz;
};

// The CPML block shared by both half-steps.  PSI selects the E-side (psiE, aE..) or H-side tables.
// dX* are the six derivative vectors of this plane; K* the curl components to correct.
// Object order of the reference: x slabs, y slabs, z slabs.
// one cell of the masked z-slab path of the aligned kernels (PM == 1)
#define FDTDX_CPML_ZCELL(E_)                                                                                           \
      if (V == 4 && (L.zm & (1 << E_))) {                                                                              \
        if (k1) cpml_axis_v<V, E_ % V, E_ % V + 1, !REV, true>(L.az, L.bz, L.kz, dzFy, dzFx, psz1, psz2, Kx, Ky);      \
        else cpml_axis_v<V, E_ % V, E_ % V + 1, !REV, false>(L.az, L.bz, L.kz, dzFy, dzFx, psz1, psz2, Kx, Ky);        \
        if (!REV && psi_st) { qz1[E_ % V] = psz1.v[E_ % V]; qz2[E_ % V] = psz2.v[E_ % V]; }                            \
      }

#define FDTDX_CPML_BLOCK(PSI, AT, BT, KT)                                                                              \
  if (PM > 0) {                                                                                                        \
    if (in_x) {                                                                                                        \
      float a = px.AT[i], b = px.BT[i];                                                                                \
      const float km1 = px.KT[i];                                                                                      \
      if (!P.simulate) { a = 0.0f; b = 1.0f; }                                                                         \
      if (px.kappa_one) cpml_axis<V, !REV, true>(a, b, km1, dxFz, dxFy, psx1, psx2, Ky, Kz);                           \
      else cpml_axis<V, !REV, false>(a, b, km1, dxFz, dxFy, psx1, psx2, Ky, Kz);                                       \
      if (!REV && psi_st) { stv<V>(qx1, psx1, nv); stv<V>(qx2, psx2, nv); }                                                    \
    }                                                                                                                  \
    if (L.in_y) {                                                                                                      \
      if (py.kappa_one) cpml_axis<V, !REV, true>(L.ay, L.by, L.ky, dyFx, dyFz, psy1, psy2, Kz, Kx);                    \
      else cpml_axis<V, !REV, false>(L.ay, L.by, L.ky, dyFx, dyFz, psy1, psy2, Kz, Kx);                                \
      if (!REV && psi_st) { stv<V>(qy1, psy1, nv); stv<V>(qy2, psy2, nv); }                                                    \
    }                                                                                                                  \
    if (PM == 2) {                                                                                                     \
      if constexpr (V == 4) {                                                                                          \
        if (L.zh0) {                                                                                                   \
          if (pz.kappa_one) cpml_axis_v<V, 0, 2, !REV, true>(L.az, L.bz, L.kz, dzFy, dzFx, psz1, psz2, Kx, Ky);        \
          else cpml_axis_v<V, 0, 2, !REV, false>(L.az, L.bz, L.kz, dzFy, dzFx, psz1, psz2, Kx, Ky);                    \
          if (!REV && psi_st) {                                                                                        \
            *reinterpret_cast<float2*>(qz1) = make_float2(psz1.v[0], psz1.v[1]);                                       \
            *reinterpret_cast<float2*>(qz2) = make_float2(psz2.v[0], psz2.v[1]);                                       \
          }                                                                                                            \
        }                                                                                                              \
        if (L.zh1) {                                                                                                   \
          if (pz.kappa_one) cpml_axis_v<V, 2, 4, !REV, true>(L.az, L.bz, L.kz, dzFy, dzFx, psz1, psz2, Kx, Ky);        \
          else cpml_axis_v<V, 2, 4, !REV, false>(L.az, L.bz, L.kz, dzFy, dzFx, psz1, psz2, Kx, Ky);                    \
          if (!REV && psi_st) {                                                                                        \
            *reinterpret_cast<float2*>(qz1 + 2) = make_float2(psz1.v[2], psz1.v[3]);                                   \
            *reinterpret_cast<float2*>(qz2 + 2) = make_float2(psz2.v[2], psz2.v[3]);                                   \
          }                                                                                                            \
        }                                                                                                              \
      }                                                                                                                \
    } else if (!FDTDX_RAGGED && L.zm) {                                                                                \
      const bool k1 = pz.kappa_one;                                                                                    \
      FDTDX_CPML_ZCELL(0) FDTDX_CPML_ZCELL(1) FDTDX_CPML_ZCELL(2) FDTDX_CPML_ZCELL(3)                                  \
    } else if (FDTDX_RAGGED && L.any_z) {                                                                              \
      _Pragma("unroll") for (int e = 0; e < V; ++e) {                                                                  \
        const int k = k0 + e * FDTDX_ES;                                                                               \
        if (k < nz && (k < pz.lo_len || k >= pz.hi_start)) {                                                           \
          const int side = (k >= pz.hi_start) ? 1 : 0;                                                                 \
          const int kl = side ? k - pz.hi_start : k;                                                                   \
          const int Lz = side ? pz.hi_len : pz.lo_len;                                                                 \
          const long long pidx = ((long long)i * ny + j) * Lz + kl;                                                    \
          float* s1 = pz.PSI[side][0] + pidx;                                                                          \
          float* s2 = pz.PSI[side][1] + pidx;                                                                          \
          float q1 = *s1, q2 = *s2;                                                                                    \
          if (!REV && P.simulate) {                                                                                    \
            q1 = pz.BT[k] * q1 + pz.AT[k] * dzFy.v[e];                                                                 \
            q2 = pz.BT[k] * q2 + pz.AT[k] * dzFx.v[e];                                                                 \
            if (psi_st) { *s1 = q1; *s2 = q2; }                                                                        \
          }                                                                                                            \
          float c1 = q1, c2 = q2;                                                                                      \
          if (!pz.kappa_one) { c1 = pz.KT[k] * dzFy.v[e] + q1; c2 = pz.KT[k] * dzFx.v[e] + q2; }                       \
          Kx.v[e] = Kx.v[e] - c1;                                                                                      \
          Ky.v[e] = Ky.v[e] + c2;                                                                                      \
        }                                                                                                              \
      }                                                                                                                \
    }                                                                                                                  \
  }

// Loads of the CPML auxiliary fields of plane i (issued together with the field loads: their
// addresses depend on indices only, so the latencies overlap).
#define FDTDX_CPML_LOADS(PSI)                                                                                          \
  const bool in_x = (PM > 0) && lane_ok && (i < px.lo_len || i >= px.hi_start);                                        \
  Vec<V> psx1, psx2, psy1, psy2, psz1, psz2;                                                                           \
  float *qx1 = nullptr, *qx2 = nullptr, *qy1 = nullptr, *qy2 = nullptr, *qz1 = nullptr, *qz2 = nullptr;                \
  if (PM > 0) {                                                                                                        \
    if (in_x) {                                                                                                        \
      const int side = (i >= px.hi_start) ? 1 : 0;                                                                     \
      const long long pidx = (long long)(side ? i - px.hi_start : i) * plane + row;                                    \
      qx1 = px.PSI[side][0] + pidx;                                                                                    \
      qx2 = px.PSI[side][1] + pidx;                                                                                    \
      psx1 = ldv<V>(qx1, nv);                                                                                              \
      psx2 = ldv<V>(qx2, nv);                                                                                              \
    }                                                                                                                  \
    if (L.in_y) {                                                                                                      \
      const long long pidx = (long long)i * L.ystride + L.yoff;                                                        \
      qy1 = py1 + pidx;                                                                                                \
      qy2 = py2 + pidx;                                                                                                \
      psy1 = ldv<V>(qy1, nv);                                                                                              \
      psy2 = ldv<V>(qy2, nv);                                                                                              \
    }                                                                                                                  \
    if (PM == 2) {                                                                                                     \
      if constexpr (V == 4) {                                                                                          \
        if (L.zvec) {                                                                                                  \
          const long long pidx = (long long)i * L.zstride + L.zoff;                                                    \
          qz1 = pz1 + pidx;                                                                                            \
          qz2 = pz2 + pidx;                                                                                            \
          if (L.zh0) {                                                                                                 \
            const float2 t1 = *reinterpret_cast<const float2*>(qz1), t2 = *reinterpret_cast<const float2*>(qz2);       \
            psz1.v[0] = t1.x; psz1.v[1] = t1.y; psz2.v[0] = t2.x; psz2.v[1] = t2.y;                                    \
          }                                                                                                            \
          if (L.zh1) {                                                                                                 \
            const float2 t1 = *reinterpret_cast<const float2*>(qz1 + 2), t2 = *reinterpret_cast<const float2*>(qz2 + 2); \
            psz1.v[2] = t1.x; psz1.v[3] = t1.y; psz2.v[2] = t2.x; psz2.v[3] = t2.y;                                    \
          }                                                                                                            \
        }                                                                                                              \
      }                                                                                                                \
    }                                                                                                                  \
    if (!FDTDX_RAGGED && PM == 1 && L.zm) {                                                                            \
      const long long pidx = (long long)i * L.zstride + L.zoff;                                                        \
      qz1 = pz1 + pidx;                                                                                                \
      qz2 = pz2 + pidx;                                                                                                \
      _Pragma("unroll") for (int e = 0; e < V; ++e)                                                                    \
        if (L.zm & (1 << e)) { psz1.v[e] = qz1[e]; psz2.v[e] = qz2[e]; }                                               \
    }                                                                                                                  \
  }

// Loop-invariant slab membership of this thread (y: warp-uniform, z: per lane).
#define FDTDX_CPML_SETUP(PSI, AT, BT, KT)                                                                              \
  const AxisPmlDev& px = P.pml[0];                                                                                     \
  const AxisPmlDev& py = P.pml[1];                                                                                     \
  const AxisPmlDev& pz = P.pml[2];                                                                                     \
  PmlLane<V> L;                                                                                                        \
  L.in_y = false; L.any_z = false; L.zvec = false; L.zh0 = false; L.zh1 = false; L.zm = 0;                             \
  float *py1 = nullptr, *py2 = nullptr, *pz1 = nullptr, *pz2 = nullptr;                                                \
  if (PM > 0 && lane_ok) {                                                                                                        \
    L.in_y = (j < py.lo_len || j >= py.hi_start);                                                                      \
    if (L.in_y) {                                                                                                      \
      const int yside = (j >= py.hi_start) ? 1 : 0;                                                                    \
      const int yL = yside ? py.hi_len : py.lo_len;                                                                    \
      L.ystride = (long long)yL * nz;                                                                                  \
      L.yoff = (long long)(yside ? j - py.hi_start : j) * nz + k0;                                                     \
      py1 = py.PSI[yside][0];                                                                                          \
      py2 = py.PSI[yside][1];                                                                                          \
      L.ay = py.AT[j]; L.by = py.BT[j]; L.ky = py.KT[j];                                                               \
      if (!P.simulate) { L.ay = 0.0f; L.by = 1.0f; }                                                                   \
    }                                                                                                                  \
    _Pragma("unroll") for (int e = 0; e < V; ++e) {                                                                    \
      const int k = k0 + e * FDTDX_ES;                                                                                 \
      if (k < nz && (k < pz.lo_len || k >= pz.hi_start)) L.any_z = true;                                               \
    }                                                                                                                  \
    if (PM == 2 && L.any_z) {                                                                                          \
      const int zside = (k0 + V > pz.hi_start) ? 1 : 0;                                                                \
      const int zL = zside ? pz.hi_len : pz.lo_len;                                                                    \
      L.zvec = true;                                                                                                   \
      L.zh0 = zside ? (k0 >= pz.hi_start) : (k0 < pz.lo_len);                                                          \
      L.zh1 = zside ? (k0 + 2 >= pz.hi_start) : (k0 + 2 < pz.lo_len);                                                  \
      L.zstride = (long long)ny * zL;                                                                                  \
      L.zoff = (long long)j * zL + (zside ? k0 - pz.hi_start : k0);                                                    \
      pz1 = pz.PSI[zside][0];                                                                                          \
      pz2 = pz.PSI[zside][1];                                                                                          \
      L.az = ldv<V>(pz.AT + k0, nv); L.bz = ldv<V>(pz.BT + k0, nv); L.kz = ldv<V>(pz.KT + k0, nv);                                 \
      if (!P.simulate) {                                                                                               \
        _Pragma("unroll") for (int e = 0; e < V; ++e) { L.az.v[e] = 0.0f; L.bz.v[e] = 1.0f; }                          \
      }                                                                                                                \
    }                                                                                                                  \
    if (!FDTDX_RAGGED && PM == 1 && L.any_z) {                                                                         \
      /* aligned rows, slabs of any thickness / position: per-cell mask, 32-bit psi accesses */                      \
      const int zside = (k0 + V > pz.hi_start) ? 1 : 0;                                                                \
      const int zL = zside ? pz.hi_len : pz.lo_len;                                                                    \
      _Pragma("unroll") for (int e = 0; e < V; ++e) {                                                                  \
        const int k = k0 + e;                                                                                          \
        if (zside ? (k >= pz.hi_start && k < nz) : (k < pz.lo_len)) L.zm |= 1 << e;                                    \
      }                                                                                                                \
      L.zstride = (long long)ny * zL;                                                                                  \
      L.zoff = (long long)j * zL + (zside ? k0 - pz.hi_start : k0);                                                    \
      pz1 = pz.PSI[zside][0];                                                                                          \
      pz2 = pz.PSI[zside][1];                                                                                          \
      L.az = ldv<V>(pz.AT + k0, nv); L.bz = ldv<V>(pz.BT + k0, nv); L.kz = ldv<V>(pz.KT + k0, nv);                     \
      if (!P.simulate) {                                                                                               \
        _Pragma("unroll") for (int e = 0; e < V; ++e) { L.az.v[e] = 0.0f; L.bz.v[e] = 1.0f; }                          \
      }                                                                                                                \
    }                                                                                                                  \
  }                                                                                                                    \
  const bool psi_st = P.simulate && P.psi_store;

// ------------------------------------------------------------------------------------------------
// Material update of one plane for the V cells of a thread (update.py:298-354, 526-609 for E;
// 736-750, 856-930 for H), shared by the marching and the TMA-staged kernels.  Conductivity, the ADE
// polarisation arrays and their coefficients move as whole vectors (128-bit, or nv predicated 32-bit
// accesses on ragged rows); `ok` is false for lanes that overhang the grid (staged kernels).
// Arithmetic order is the reference's, element by element.
// ------------------------------------------------------------------------------------------------
// conductivity vectors of one plane (1 or 3 components): issued with the field loads so that their
// latency overlaps the tile wait / curl instead of stalling the material update
template <int V>
__device__ __forceinline__ void load_sigma(const float* sig, const long long cs, const long long cell0, const bool ok, const int nv, Vec<V> (&sg)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c == 0 || cs != 0) sg[c] = ok ? ldv<V>(sig + c * cs + cell0, nv) : zerov<V>();
    else sg[c] = sg[0];
  }
}
// L2 prefetch of the ADE arrays of a later plane (polarisations and coefficients of every pole / component)
__device__ __forceinline__ void prefetch_ade(const StepParams& P, const long long N, const long long cell) {
  const long long cstride = P.c_cs ? 3 * N : N;
  for (int p = 0; p < P.n_poles; ++p) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long pi = p * 3 * N + c * N + cell;
      prefetch_l2(P.P_cur + pi);
      prefetch_l2(P.P_new + pi);
      if (c == 0 || P.c_cs != 0) {
        const long long ci = p * cstride + c * P.c_cs + cell;
        prefetch_l2(P.c1 + ci); prefetch_l2(P.c2 + ci); prefetch_l2(P.c3 + ci);
        if (P.has_c4) prefetch_l2(P.c4 + ci);
      }
    }
  }
}

template <int V, bool REV, bool SIG, bool ADE>
__device__ __forceinline__ void material_update_E(const StepParams& P, const long long N, const long long cell0, const bool ok, const int nv,
                                                  const Vec<V> (&Eo)[3], const Vec<V> (&K)[3], const Vec<V> (&ie)[3], const Vec<V> (&sg)[3],
                                                  Vec<V> (&out)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Vec<V> s, E1;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      if (REV) {
        // update_E_reverse: ((1+s)E - c K inv_eps) / (1-s)
        if (SIG) {
          const float sv = (((P.cour * sg[c].v[e]) * P.eta0) * ie[c].v[e]) / 2.0f;
          const float Ec = Eo[c].v[e] * (1.0f + sv);
          E1.v[e] = (Ec - (P.cour * K[c].v[e]) * ie[c].v[e]) / (1.0f - sv);
        } else {
          E1.v[e] = Eo[c].v[e] - (P.cour * K[c].v[e]) * ie[c].v[e];
        }
      } else if (SIG) {
        s.v[e] = (((P.cour * sg[c].v[e]) * P.eta0) * ie[c].v[e]) / 2.0f;
        E1.v[e] = (1.0f - s.v[e]) * Eo[c].v[e] + (P.cour * K[c].v[e]) * ie[c].v[e];
      } else {
        s.v[e] = 0.0f;
        E1.v[e] = Eo[c].v[e] + (P.cour * K[c].v[e]) * ie[c].v[e];
      }
    }
    if (!REV && ADE) {
      if (ok) {
        // P_hat = c1 P + c2 P_prev + c3 E ; E += inv_eps * sum_p (P - P_hat)   (update.py:330-332)
        const long long pstride = 3 * N;
        const long long cstride = P.c_cs ? 3 * N : N;
        Vec<V> delta = zerov<V>(), c4sum = zerov<V>();
        for (int p = 0; p < P.n_poles; ++p) {
          const long long pi = p * pstride + c * N + cell0;
          const long long ci = p * cstride + c * P.c_cs + cell0;
          const Vec<V> Pc = ldv<V>(P.P_cur + pi, nv), Pp = ldv<V>(P.P_new + pi, nv);
          const Vec<V> a1 = ldv<V>(P.c1 + ci, nv), a2 = ldv<V>(P.c2 + ci, nv), a3 = ldv<V>(P.c3 + ci, nv);
          Vec<V> a4 = zerov<V>();
          if (P.has_c4) a4 = ldv<V>(P.c4 + ci, nv);
          Vec<V> Phat;
#pragma unroll
          for (int e = 0; e < V; ++e) {
            Phat.v[e] = (a1.v[e] * Pc.v[e] + a2.v[e] * Pp.v[e]) + a3.v[e] * Eo[c].v[e];
            const float dd = Pc.v[e] - Phat.v[e];
            delta.v[e] = (p == 0) ? dd : delta.v[e] + dd;
            if (P.has_c4) c4sum.v[e] = (p == 0) ? a4.v[e] : c4sum.v[e] + a4.v[e];
          }
          if (P.p_store) stv<V>(P.P_new + pi, Phat, nv);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
          E1.v[e] = E1.v[e] + ie[c].v[e] * delta.v[e];
          if (P.has_c4) {
            float divisor = 1.0f + ie[c].v[e] * c4sum.v[e];
            if (SIG) divisor = divisor + s.v[e];
            E1.v[e] = E1.v[e] / divisor;
          } else if (SIG) {
            E1.v[e] = E1.v[e] / (1.0f + s.v[e]);
          }
        }
        if (P.has_c4 && P.p_store) {
          for (int p = 0; p < P.n_poles; ++p) {
            const long long pi = p * pstride + c * N + cell0;
            const long long ci = p * cstride + c * P.c_cs + cell0;
            Vec<V> Pn = ldv<V>(P.P_new + pi, nv);
            const Vec<V> a4 = ldv<V>(P.c4 + ci, nv);
#pragma unroll
            for (int e = 0; e < V; ++e) Pn.v[e] = Pn.v[e] + a4.v[e] * E1.v[e];
            stv<V>(P.P_new + pi, Pn, nv);
          }
        }
      }
    } else if (!REV && SIG) {
#pragma unroll
      for (int e = 0; e < V; ++e) E1.v[e] = E1.v[e] / (1.0f + s.v[e]);
    }
    out[c] = E1;
  }
}

template <int V, bool REV, bool SIG>
__device__ __forceinline__ void material_update_H(const StepParams& P, const long long cell0, const bool ok, const int nv,
                                                  const Vec<V> (&Ho)[3], const Vec<V> (&K)[3], const Vec<V> (&im)[3], const Vec<V> (&sg)[3],
                                                  Vec<V> (&out)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      if (SIG) {
        const float sv = (((P.cour * sg[c].v[e]) / P.eta0) * im[c].v[e]) / 2.0f;
        if (REV) {
          const float Hc = Ho[c].v[e] * (1.0f + sv);
          out[c].v[e] = (Hc + (P.cour * K[c].v[e]) * im[c].v[e]) / (1.0f - sv);
        } else {
          const float H1 = (1.0f - sv) * Ho[c].v[e] - (P.cour * K[c].v[e]) * im[c].v[e];
          out[c].v[e] = H1 / (1.0f + sv);
        }
      } else if (REV) {
        out[c].v[e] = Ho[c].v[e] + (P.cour * K[c].v[e]) * im[c].v[e];
      } else {
        out[c].v[e] = Ho[c].v[e] - (P.cour * K[c].v[e]) * im[c].v[e];
      }
    }
  }
}

#if !defined(FDTDX_BUILD_H)
// ------------------------------------------------------------------------------------------------
// E half-step
// ------------------------------------------------------------------------------------------------
// KONLY: write the curl (with its CPML correction) to P.E instead of updating it - phase 1 of the
// full-tensor tier (tensor_kernels.cuh); E, inv_eps, sources and walls are not touched.
template <int V, int TIER, bool REV, bool SIG, bool ADE, bool MET, int PM, bool KONLY = false>
static __global__ void __launch_bounds__(256, FDTDX_MIN_CTAS) yee_E_kernel(const __grid_constant__ StepParams P, const int t) {
  const int lane = threadIdx.x;
  const int k0 = FDTDX_RAGGED ? (int)blockIdx.x * 32 * V + lane : ((int)blockIdx.x * 32 + lane) * V;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int nz = P.nz, ny = P.ny;
  const int ic0 = P.x_begin + (P.z_reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z) * P.xchunk;
  const int ic1 = min(ic0 + P.xchunk, P.x_end);
  const bool peer_cta = (ic0 == 0);  // the chunk that reads the low neighbour's H plane and owns E[0]
  if (P.peer_wait != nullptr && peer_cta) peer_wait_cta(P);
  const bool active = (k0 < nz) && (j < ny);
  const unsigned wmask = __ballot_sync(0xffffffffu, active);
  if (!active) return;  // the shuffles below run under wmask
  const int nv = FDTDX_RAGGED ? min(V, (nz - k0 + FDTDX_ES - 1) / FDTDX_ES) : V;  // valid elements of this thread
  const long long plane = (long long)ny * nz;
  const long long N = plane * P.nx;
  const long long row = (long long)j * nz + k0;

  // neighbour offsets (in elements) relative to this thread's first cell; *_ok false = zero halo
  bool jm_ok = true, km_ok = true;
  long long djm = -(long long)nz;
  if (j == 0) { if (P.wrap[1] && !P.sym[1]) djm = (long long)(ny - 1) * nz; else jm_ok = false; }
  int dkm = -1;
  if (k0 == 0) { if (P.wrap[2] && !P.sym[2]) dkm = nz - 1; else km_ok = false; }

  const bool lane_ok = true;  // inactive lanes have already left
  FDTDX_CPML_SETUP(psiE, aE, bE, kE)

  float sBy = 1.0f;
  Vec<V> sBz;
  if (MET) {
    sBy = P.sB[1][j];
    sBz = ldv<V>(P.sB[2] + k0, nv);
  }

  const float* pH = P.H + (long long)ic0 * plane + row;  // Hx of (ic0, j, k0); Hy, Hz at +N, +2N
  float* pE = P.E + (long long)ic0 * plane + row;
  const float* pEps = P.eps + (long long)ic0 * plane + row;

  // register queue: Hy, Hz of the previous x plane
  Vec<V> hy_im, hz_im;
  if (ic0 > 0) {
    hy_im = ldv<V>(pH + N - plane, nv);
    hz_im = ldv<V>(pH + 2 * N - plane, nv);
  } else if (P.x_lo_mode == 1) {
    const long long o = (long long)(P.nx - 1) * plane + row;
    hy_im = ldv<V>(P.H + N + o, nv);
    hz_im = ldv<V>(P.H + 2 * N + o, nv);
    if (P.bH != nullptr) {  // Bloch ghost plane: H[nx-1] * conj(phase_x)
      hy_im = bloch_mix<V>(hy_im, ldv<V>(P.bH + N + o, nv), P.bc[0], P.bs[0]);
      hz_im = bloch_mix<V>(hz_im, ldv<V>(P.bH + 2 * N + o, nv), P.bc[0], P.bs[0]);
    }
  } else if (P.x_lo_mode == 2) {
    hy_im = ldv<V>(P.haloH + row, nv);
    hz_im = ldv<V>(P.haloH + P.haloH_cs + row, nv);
  } else {
    hy_im = zerov<V>();
    hz_im = zerov<V>();
  }

  if (!KONLY && REV && P.n_src > 0 && P.src_inline && ic0 < P.src_x1 && ic1 > P.src_x0) src_pass_E<V, TIER>(P, t, true, ic0, ic1, j, k0);
  for (int i = ic0; i < ic1; ++i) {
    const Vec<V> hx = ldv<V>(pH, nv), hy = ldv<V>(pH + N, nv), hz = ldv<V>(pH + 2 * N, nv);
    Vec<V> hx_jm, hz_jm;
    if (jm_ok) {
      hx_jm = ldv<V>(pH + djm, nv);
      hz_jm = ldv<V>(pH + 2 * N + djm, nv);
      if (P.bH != nullptr && j == 0) {  // Bloch ghost row: H[ny-1] * conj(phase_y)
        const float* q = P.bH + (pH - P.H) + djm;
        hx_jm = bloch_mix<V>(hx_jm, ldv<V>(q, nv), P.bc[1], P.bs[1]);
        hz_jm = bloch_mix<V>(hz_jm, ldv<V>(q + 2 * N, nv), P.bc[1], P.bs[1]);
      }
    } else {
      hx_jm = zerov<V>();
      hz_jm = zerov<V>();
    }
    Vec<V> ex, ey, ez, ie0, ie1, ie2, sg3[3];
    if (SIG) load_sigma<V>(P.sigE, P.sigE_cs, pE - P.E, true, nv, sg3);
    if (!KONLY) {
      ex = ldv<V>(pE, nv); ey = ldv<V>(pE + N, nv); ez = ldv<V>(pE + 2 * N, nv);
      ie0 = ldv<V>(pEps, nv);
      if (TIER == 3) {
        ie1 = ldv<V>(pEps + P.eps_cs, nv);
        ie2 = ldv<V>(pEps + 2 * P.eps_cs, nv);
      } else {
        ie1 = ie0;
        ie2 = ie0;
      }
    }
    FDTDX_CPML_LOADS(psiE)
    // L2 prefetch of this thread's lines FDTDX_PF_DIST planes ahead (holds no registers)
    if (i + FDTDX_PF_DIST < ic1) {
      const long long pb = FDTDX_PF_DIST * plane;
      prefetch_l2(pH + pb); prefetch_l2(pH + N + pb); prefetch_l2(pH + 2 * N + pb);
      if (SIG) { prefetch_l2(P.sigE + (pE - P.E) + pb); if (P.sigE_cs) { prefetch_l2(P.sigE + P.sigE_cs + (pE - P.E) + pb); prefetch_l2(P.sigE + 2 * P.sigE_cs + (pE - P.E) + pb); } }
      if (ADE) prefetch_ade(P, N, (pE - P.E) + pb);
      if (!KONLY) {
        prefetch_l2(pE + pb); prefetch_l2(pE + N + pb); prefetch_l2(pE + 2 * N + pb);
        prefetch_l2(pEps + pb);
        if (TIER == 3) { prefetch_l2(pEps + P.eps_cs + pb); prefetch_l2(pEps + 2 * P.eps_cs + pb); }
      }
    }
    // z-neighbour (k-1) of the first element: last element of the previous lane
    Vec<V> hx_kmv, hy_kmv;
    if constexpr (FDTDX_RAGGED) {
      // interleaved cells: k-1 of (element e, lane l) is (e, l-1); for lane 0 it is (e-1, lane 31)
#pragma unroll
      for (int e = 0; e < V; ++e) {
        hx_kmv.v[e] = __shfl_up_sync(wmask, hx.v[e], 1);
        hy_kmv.v[e] = __shfl_up_sync(wmask, hy.v[e], 1);
        if (e > 0) {
          const float tx = __shfl_sync(wmask, hx.v[e - 1], 31), ty = __shfl_sync(wmask, hy.v[e - 1], 31);
          if (lane == 0) { hx_kmv.v[e] = tx; hy_kmv.v[e] = ty; }
        }
      }
      if (lane == 0) {
        hx_kmv.v[0] = km_ok ? pH[dkm] : 0.0f;
        hy_kmv.v[0] = km_ok ? pH[N + dkm] : 0.0f;
        if (P.bH != nullptr && k0 == 0 && km_ok) {  // Bloch ghost column: H[nz-1] * conj(phase_z)
          const float* q = P.bH + (pH - P.H) + dkm;
          hx_kmv.v[0] = hx_kmv.v[0] * P.bc[2] + q[0] * P.bs[2];
          hy_kmv.v[0] = hy_kmv.v[0] * P.bc[2] + q[N] * P.bs[2];
        }
      }
    } else {
      float hx_l = __shfl_up_sync(wmask, hx.v[V - 1], 1);
      float hy_l = __shfl_up_sync(wmask, hy.v[V - 1], 1);
      if (lane == 0) {
        hx_l = km_ok ? pH[dkm] : 0.0f;
        hy_l = km_ok ? pH[N + dkm] : 0.0f;
        if (P.bH != nullptr && k0 == 0 && km_ok) {  // Bloch ghost column: H[nz-1] * conj(phase_z)
          const float* q = P.bH + (pH - P.H) + dkm;
          hx_l = hx_l * P.bc[2] + q[0] * P.bs[2];
          hy_l = hy_l * P.bc[2] + q[N] * P.bs[2];
        }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        hx_kmv.v[e] = (e == 0) ? hx_l : hx.v[e == 0 ? 0 : e - 1];
        hy_kmv.v[e] = (e == 0) ? hy_l : hy.v[e == 0 ? 0 : e - 1];
      }
    }
    float sBx = 1.0f;
    if (MET) sBx = P.sB[0][i];
    Vec<V> Kx, Ky, Kz;
    Vec<V> dxFz, dxFy, dyFx, dyFz, dzFy, dzFx;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float hx_km = hx_kmv.v[e], hy_km = hy_kmv.v[e];
      float dyHz = hz.v[e] - hz_jm.v[e];
      float dzHy = hy.v[e] - hy_km;
      float dzHx = hx.v[e] - hx_km;
      float dxHz = hz.v[e] - hz_im.v[e];
      float dxHy = hy.v[e] - hy_im.v[e];
      float dyHx = hx.v[e] - hx_jm.v[e];
      if (MET) {
        dyHz *= sBy; dzHy *= sBz.v[e]; dzHx *= sBz.v[e]; dxHz *= sBx; dxHy *= sBx; dyHx *= sBy;
      }
      Kx.v[e] = dyHz - dzHy;
      Ky.v[e] = dzHx - dxHz;
      Kz.v[e] = dxHy - dyHx;
      dxFz.v[e] = dxHz; dxFy.v[e] = dxHy; dyFx.v[e] = dyHx;
      dyFz.v[e] = dyHz; dzFy.v[e] = dzHy; dzFx.v[e] = dzHx;
    }
    FDTDX_CPML_BLOCK(psiE, aE, bE, kE)
    // material update
    Vec<V> o3[3];
    if (KONLY) {
      o3[0] = Kx; o3[1] = Ky; o3[2] = Kz;
    } else {
      const Vec<V> Eo3[3] = {ex, ey, ez}, K3[3] = {Kx, Ky, Kz}, ie3[3] = {ie0, ie1, ie2};
      material_update_E<V, REV, SIG, ADE>(P, N, pE - P.E, true, nv, Eo3, K3, ie3, sg3, o3);
    }
    Vec<V>&o0 = o3[0], &o1 = o3[1], &o2 = o3[2];
    // PEC walls (pec.py:70-77)
    if (!KONLY && P.n_walls > 0 && i >= P.wall_x0[0] && i < P.wall_x1[0]) wall_mask<V>(P, 0, i, j, k0, o0, o1, o2);
    stv<V>(pE, o0, nv);
    stv<V>(pE + N, o1, nv);
    stv<V>(pE + 2 * N, o2, nv);
    hy_im = hy;
    hz_im = hz;
    pH += plane;
    pE += plane;
    pEps += plane;
  }
  if (!KONLY && !REV && P.n_src > 0 && P.src_inline && ic0 < P.src_x1 && ic1 > P.src_x0) src_pass_E<V, TIER>(P, t, false, ic0, ic1, j, k0);
  if (P.peer_signal != nullptr && peer_cta) peer_signal_warp(P, wmask);
}

#endif  // !FDTDX_BUILD_H

#if !defined(FDTDX_BUILD_E)
// ------------------------------------------------------------------------------------------------
// H half-step
// ------------------------------------------------------------------------------------------------
template <int V, int MUT, bool REV, bool SIG, bool MET, int PM, bool KONLY = false>
static __global__ void __launch_bounds__(256, FDTDX_MIN_CTAS) yee_H_kernel(const __grid_constant__ StepParams P, const int t) {
  const int lane = threadIdx.x;
  const int k0 = FDTDX_RAGGED ? (int)blockIdx.x * 32 * V + lane : ((int)blockIdx.x * 32 + lane) * V;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int nz = P.nz, ny = P.ny;
  const int ic0 = P.x_begin + blockIdx.z * P.xchunk;
  const int ic1 = min(ic0 + P.xchunk, P.x_end);
  const bool peer_cta = (ic1 == P.nx);  // the chunk that reads the high neighbour's E plane and owns H[nx-1]
  if (P.peer_wait != nullptr && peer_cta) peer_wait_cta(P);
  const bool active = (k0 < nz) && (j < ny);
  const unsigned wmask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const int nv = FDTDX_RAGGED ? min(V, (nz - k0 + FDTDX_ES - 1) / FDTDX_ES) : V;
  const long long plane = (long long)ny * nz;
  const long long N = plane * P.nx;
  const long long row = (long long)j * nz + k0;

  bool jp_ok = true, kp_ok = true;
  long long djp = nz;
  if (j == ny - 1) { if (P.wrap[1]) djp = -(long long)(ny - 1) * nz; else jp_ok = false; }
  int dkp = V;  // first cell after this thread's vector
  if (k0 + V >= nz) { if (P.wrap[2]) dkp = -k0; else kp_ok = false; }
  const bool last_lane = (lane == 31) || (k0 + V >= nz);

  const bool lane_ok = true;
  FDTDX_CPML_SETUP(psiH, aH, bH, kH)

  float sFy = 1.0f;
  Vec<V> sFz;
  if (MET) {
    sFy = P.sF[1][j];
    sFz = ldv<V>(P.sF[2] + k0, nv);
  }

  const float* pE = P.E + (long long)ic0 * plane + row;
  float* pH = P.H + (long long)ic0 * plane + row;
  const float* pMu = (MUT >= 1) ? P.mu + (long long)ic0 * plane + row : nullptr;

  // register queue: E of the current plane is the "next" plane loaded one step earlier
  Vec<V> ex = ldv<V>(pE, nv), ey = ldv<V>(pE + N, nv), ez = ldv<V>(pE + 2 * N, nv);

  if (!KONLY && REV && P.n_src > 0 && P.src_inline && ic0 < P.src_x1 && ic1 > P.src_x0) src_pass_H<V, MUT>(P, t, true, ic0, ic1, j, k0);
  for (int i = ic0; i < ic1; ++i) {
    Vec<V> ex_n, ey_n, ez_n;
    if (i + 1 < P.nx) {
      ey_n = ldv<V>(pE + N + plane, nv);
      ez_n = ldv<V>(pE + 2 * N + plane, nv);
      if (i + 1 < ic1) ex_n = ldv<V>(pE + plane, nv);
    } else if (P.x_hi_mode == 1) {
      ey_n = ldv<V>(P.E + N + row, nv);
      ez_n = ldv<V>(P.E + 2 * N + row, nv);
      if (P.bE != nullptr) {  // Bloch ghost plane: E[0] * phase_x
        ey_n = bloch_mix<V>(ey_n, ldv<V>(P.bE + N + row, nv), P.bc[0], -P.bs[0]);
        ez_n = bloch_mix<V>(ez_n, ldv<V>(P.bE + 2 * N + row, nv), P.bc[0], -P.bs[0]);
      }
    } else if (P.x_hi_mode == 2) {
      ey_n = ldv<V>(P.haloE + row, nv);
      ez_n = ldv<V>(P.haloE + P.haloE_cs + row, nv);
    } else {
      ey_n = zerov<V>();
      ez_n = zerov<V>();
    }
    Vec<V> ex_jp, ez_jp;
    if (jp_ok) {
      ex_jp = ldv<V>(pE + djp, nv);
      ez_jp = ldv<V>(pE + 2 * N + djp, nv);
      if (P.bE != nullptr && j == ny - 1) {  // Bloch ghost row: E[0] * phase_y
        const float* q = P.bE + (pE - P.E) + djp;
        ex_jp = bloch_mix<V>(ex_jp, ldv<V>(q, nv), P.bc[1], -P.bs[1]);
        ez_jp = bloch_mix<V>(ez_jp, ldv<V>(q + 2 * N, nv), P.bc[1], -P.bs[1]);
      }
    } else {
      ex_jp = zerov<V>();
      ez_jp = zerov<V>();
    }
    Vec<V> hx, hy, hz, sg3[3];
    if (SIG) load_sigma<V>(P.sigH, P.sigH_cs, pH - P.H, true, nv, sg3);
    if (!KONLY) { hx = ldv<V>(pH, nv); hy = ldv<V>(pH + N, nv); hz = ldv<V>(pH + 2 * N, nv); }
    Vec<V> im0, im1, im2;
    if (!KONLY && MUT >= 1) {
      im0 = ldv<V>(pMu, nv);
      if (MUT == 3) {
        im1 = ldv<V>(pMu + P.mu_cs, nv);
        im2 = ldv<V>(pMu + 2 * P.mu_cs, nv);
      } else {
        im1 = im0;
        im2 = im0;
      }
    }
    FDTDX_CPML_LOADS(psiH)
    if (i + FDTDX_PF_DIST < ic1) {
      const long long pb = FDTDX_PF_DIST * plane;
      if (SIG) { prefetch_l2(P.sigH + (pH - P.H) + pb); if (P.sigH_cs) { prefetch_l2(P.sigH + P.sigH_cs + (pH - P.H) + pb); prefetch_l2(P.sigH + 2 * P.sigH_cs + (pH - P.H) + pb); } }
      if (!KONLY) { prefetch_l2(pH + pb); prefetch_l2(pH + N + pb); prefetch_l2(pH + 2 * N + pb); }
      prefetch_l2(pE + pb); prefetch_l2(pE + N + pb); prefetch_l2(pE + 2 * N + pb);
      if (!KONLY && MUT >= 1) prefetch_l2(pMu + pb);
      if (MUT == 3) { prefetch_l2(pMu + P.mu_cs + pb); prefetch_l2(pMu + 2 * P.mu_cs + pb); }
    }
    Vec<V> ex_kpv, ey_kpv;
    if constexpr (FDTDX_RAGGED) {
      // interleaved cells: k+1 of (element e, lane l) is (e, l+1); for lane 31 it is (e+1, lane 0);
      // past the row end it is the z halo (zero, or the row's first cell on a periodic axis)
#pragma unroll
      for (int e = 0; e < V; ++e) {
        ex_kpv.v[e] = __shfl_down_sync(wmask, ex.v[e], 1);
        ey_kpv.v[e] = __shfl_down_sync(wmask, ey.v[e], 1);
        if (e + 1 < V) {
          const float tx = __shfl_sync(wmask, ex.v[e + 1], 0), ty = __shfl_sync(wmask, ey.v[e + 1], 0);
          if (lane == 31) { ex_kpv.v[e] = tx; ey_kpv.v[e] = ty; }
        }
        const int kn = k0 + e * FDTDX_ES + 1;
        if (kn >= nz) {
          ex_kpv.v[e] = P.wrap[2] ? pE[-k0] : 0.0f;
          ey_kpv.v[e] = P.wrap[2] ? pE[N - k0] : 0.0f;
          if (P.bE != nullptr && P.wrap[2]) {  // Bloch ghost column: E[0] * phase_z
            const float* q = P.bE + (pE - P.E) - k0;
            ex_kpv.v[e] = ex_kpv.v[e] * P.bc[2] - q[0] * P.bs[2];
            ey_kpv.v[e] = ey_kpv.v[e] * P.bc[2] - q[N] * P.bs[2];
          }
        } else if (e == V - 1 && lane == 31) {
          ex_kpv.v[e] = pE[kn - k0];
          ey_kpv.v[e] = pE[N + kn - k0];
        }
      }
    } else {
      float ex_r = __shfl_down_sync(wmask, ex.v[0], 1);
      float ey_r = __shfl_down_sync(wmask, ey.v[0], 1);
      if (last_lane) {
        ex_r = kp_ok ? pE[dkp] : 0.0f;
        ey_r = kp_ok ? pE[N + dkp] : 0.0f;
        if (P.bE != nullptr && k0 + V >= nz && kp_ok) {  // Bloch ghost column: E[0] * phase_z
          const float* q = P.bE + (pE - P.E) + dkp;
          ex_r = ex_r * P.bc[2] - q[0] * P.bs[2];
          ey_r = ey_r * P.bc[2] - q[N] * P.bs[2];
        }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        ex_kpv.v[e] = (e == V - 1) ? ex_r : ex.v[e == V - 1 ? e : e + 1];
        ey_kpv.v[e] = (e == V - 1) ? ey_r : ey.v[e == V - 1 ? e : e + 1];
      }
    }
    float sFx = 1.0f;
    if (MET) sFx = P.sF[0][i];
    Vec<V> Kx, Ky, Kz;
    Vec<V> dxFz, dxFy, dyFx, dyFz, dzFy, dzFx;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float ex_kp = ex_kpv.v[e], ey_kp = ey_kpv.v[e];
      float dyEz = ez_jp.v[e] - ez.v[e];
      float dzEy = ey_kp - ey.v[e];
      float dzEx = ex_kp - ex.v[e];
      float dxEz = ez_n.v[e] - ez.v[e];
      float dxEy = ey_n.v[e] - ey.v[e];
      float dyEx = ex_jp.v[e] - ex.v[e];
      if (MET) {
        dyEz *= sFy; dzEy *= sFz.v[e]; dzEx *= sFz.v[e]; dxEz *= sFx; dxEy *= sFx; dyEx *= sFy;
      }
      Kx.v[e] = dyEz - dzEy;
      Ky.v[e] = dzEx - dxEz;
      Kz.v[e] = dxEy - dyEx;
      dxFz.v[e] = dxEz; dxFy.v[e] = dxEy; dyFx.v[e] = dyEx;
      dyFz.v[e] = dyEz; dzFy.v[e] = dzEy; dzFx.v[e] = dzEx;
    }
    FDTDX_CPML_BLOCK(psiH, aH, bH, kH)
    Vec<V> o3[3];
    if (KONLY) {
      o3[0] = Kx; o3[1] = Ky; o3[2] = Kz;
    } else {
      Vec<V> im3[3];
      if (MUT >= 1) { im3[0] = im0; im3[1] = im1; im3[2] = im2; }
      else {
#pragma unroll
        for (int e = 0; e < V; ++e) im3[0].v[e] = P.inv_mu_scalar;
        im3[1] = im3[0]; im3[2] = im3[0];
      }
      const Vec<V> Ho3[3] = {hx, hy, hz}, K3[3] = {Kx, Ky, Kz};
      material_update_H<V, REV, SIG>(P, pH - P.H, true, nv, Ho3, K3, im3, sg3, o3);
    }
    Vec<V>&o0 = o3[0], &o1 = o3[1], &o2 = o3[2];
    if (!KONLY && P.n_walls > 0 && i >= P.wall_x0[1] && i < P.wall_x1[1]) wall_mask<V>(P, 1, i, j, k0, o0, o1, o2);
    if (!KONLY && !REV && P.hprev_out != nullptr && hprev_wanted(P, i, j)) {  // H_prev for the detector pass
      float* hp = P.hprev_out + (pH - P.H);
      stv<V>(hp, hx, nv);
      stv<V>(hp + N, hy, nv);
      stv<V>(hp + 2 * N, hz, nv);
    }
    stv<V>(pH, o0, nv);
    stv<V>(pH + N, o1, nv);
    stv<V>(pH + 2 * N, o2, nv);
    ex = ex_n;
    ey = ey_n;
    ez = ez_n;
    pE += plane;
    pH += plane;
    if (MUT >= 1) pMu += plane;
  }
  if (!KONLY && !REV && P.n_src > 0 && P.src_inline && ic0 < P.src_x1 && ic1 > P.src_x0) src_pass_H<V, MUT>(P, t, false, ic0, ic1, j, k0);
  if (P.peer_signal != nullptr && peer_cta) peer_signal_warp(P, wmask);
}
#endif  // !FDTDX_BUILD_E
