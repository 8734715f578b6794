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
static __device__ __noinline__ float src_profile(const SrcDev& S, float time) {
  if (S.profile_kind == 0) {  // SingleFrequencyProfile
    float phase = ((S.p[4] * time) / S.p[0] + S.p[1]) + S.p[2];  // 2*pi*time/period + phase_shift + self.phase_shift
    float raw = cosf(phase);                                     // Re(exp(-i*phase))
    float f = time / S.p[3];
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    return f * raw;
  } else if (S.profile_kind == 1) {  // GaussianPulseProfile
    float d = time - S.p[3];
    float env = expf(-(d * d) / S.p[5]);
    float phase = (S.p[0] * time + S.p[1]) + S.p[2];
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
      float Hb = S.Hinc[b * fn + f] * amp_b;
      float Ha = S.Hinc[a * fn + f] * amp_a;
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
      float amp_a = src_profile(S, (tf + S.toffE[a * fn + f]) * dt) * S.static_amp;
      float amp_b = src_profile(S, (tf + S.toffE[b * fn + f]) * dt) * S.static_amp;
      float Ea = S.Einc[a * fn + f] * amp_a;
      float Eb = S.Einc[b * fn + f] * amp_b;
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

__device__ __forceinline__ bool any_src_hits(const SrcDev* __restrict__ srcs, int n_src, int i, int j, int k0, int V) {
  for (int s = 0; s < n_src; ++s) {
    const SrcDev& S = srcs[s];
    if (i >= S.lo[0] && i < S.hi[0] && j >= S.lo[1] && j < S.hi[1] && k0 < S.hi[2] && k0 + V > S.lo[2]) return true;
  }
  return false;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
#ifndef FDTDX_PF_DIST
#define FDTDX_PF_DIST 2
#endif
#ifndef FDTDX_PF_L1
#define FDTDX_PF_L1 0  // 1: additionally pull the next plane into L1
#endif

// CPML for one axis at one cell (perfectly_matched_layer.py:138-190; curl.py:284-308, 371-394).
// d1 = d_a F_j, d2 = d_a F_i; returns the corrections to subtract from K_i and add to K_j.
__device__ __forceinline__ void cpml_cell(float a, float b, float km1, bool kappa_one, bool simulate,
                                          float d1, float d2, float* psi1, float* psi2,
                                          float* corr1, float* corr2) {
  float p1 = *psi1, p2 = *psi2;
  if (simulate) {
    p1 = b * p1 + a * d1;
    p2 = b * p2 + a * d2;
    *psi1 = p1;
    *psi2 = p2;
  }
  if (kappa_one) {
    *corr1 = p1;
    *corr2 = p2;
  } else {
    *corr1 = km1 * d1 + p1;
    *corr2 = km1 * d2 + p2;
  }
}

#if !defined(FDTDX_BUILD_H)
// ------------------------------------------------------------------------------------------------
// E half-step
// ------------------------------------------------------------------------------------------------
template <int V, int TIER, bool REV, bool SIG, bool ADE, bool MET, int PM>
__global__ void __launch_bounds__(256) yee_E_kernel(const StepParams P, const int t) {
  const int lane = threadIdx.x;
  const int k0 = (blockIdx.x * 32 + lane) * V;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int ic0 = P.x_begin + blockIdx.z * P.xchunk;
  const int ic1 = min(ic0 + P.xchunk, P.x_end);
  const bool active = (k0 < P.nz) && (j < P.ny);
  const int nz = P.nz, ny = P.ny;
  const long long plane = (long long)ny * nz;
  const long long N = plane * P.nx;
  const float* __restrict__ Hx = P.H;
  const float* __restrict__ Hy = P.H + N;
  const float* __restrict__ Hz = P.H + 2 * N;
  float* __restrict__ Ex = P.E;
  float* __restrict__ Ey = P.E + N;
  float* __restrict__ Ez = P.E + 2 * N;

  const long long row = (long long)j * nz + k0;
  int jm = j - 1;
  bool jm_ok = true;
  if (jm < 0) { if (P.wrap[1]) jm = ny - 1; else jm_ok = false; }
  const long long rowm = (long long)jm * nz + k0;
  int km = k0 - 1;
  bool km_ok = true;
  if (km < 0) { if (P.wrap[2]) km = nz - 1; else km_ok = false; }
  const long long rowk = (long long)j * nz + km;

  // y / z PML membership of this thread's cells (x membership is per marching step)
  const AxisPmlDev& px = P.pml[0];
  const AxisPmlDev& py = P.pml[1];
  const AxisPmlDev& pz = P.pml[2];
  // PM: 0 = no CPML slab on this plan (all CPML code compiled out), 1 = scalar z-slab path, 2 = 128-bit z-slab path
  const bool in_y = (PM > 0) && active && (j < py.lo_len || j >= py.hi_start);
  const int yside = (j >= py.hi_start) ? 1 : 0;
  const int jl = yside ? j - py.hi_start : j;
  const int yL = yside ? py.hi_len : py.lo_len;
  const bool any_z = (PM > 0) && active && (k0 < pz.lo_len || k0 + V > pz.hi_start);
  // PM == 2 (V == 4, even slab thickness): the lane's four cells are two 64-bit halves, each either
  // entirely inside or entirely outside a z slab, so psi moves as float2 with no per-cell branch.
  const int zside = (k0 + V > pz.hi_start) ? 1 : 0;
  const bool zvec = (PM == 2) && any_z;
  const bool zh0 = zvec && (zside ? (k0 >= pz.hi_start) : (k0 < pz.lo_len));
  const bool zh1 = zvec && (zside ? (k0 + 2 >= pz.hi_start) : (k0 + 2 < pz.lo_len));
  const long long zstride = (long long)ny * (zside ? pz.hi_len : pz.lo_len);
  const long long zoff = (long long)j * (zside ? pz.hi_len : pz.lo_len) + (zside ? k0 - pz.hi_start : k0);
  float* const pz1 = zside ? pz.psiE[1][0] : pz.psiE[0][0];
  float* const pz2 = zside ? pz.psiE[1][1] : pz.psiE[0][1];
  Vec<V> az = zerov<V>(), bz = zerov<V>(), kz = zerov<V>();
  if (zvec) { az = ldv<V>(pz.aE + k0); bz = ldv<V>(pz.bE + k0); kz = ldv<V>(pz.kE + k0); }
  float sBy = 1.0f;
  float sBzv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) sBzv[e] = 1.0f;
  if (MET && active) {
    sBy = P.sB[1][j];
#pragma unroll
    for (int e = 0; e < V; ++e) sBzv[e] = P.sB[2][k0 + e];
  }

  // register queue: Hy, Hz of the previous x plane
  Vec<V> hy_im = zerov<V>(), hz_im = zerov<V>();
  if (active) {
    if (ic0 > 0) {
      hy_im = ldv<V>(Hy + (long long)(ic0 - 1) * plane + row);
      hz_im = ldv<V>(Hz + (long long)(ic0 - 1) * plane + row);
    } else if (P.x_lo_mode == 1) {
      hy_im = ldv<V>(Hy + (long long)(P.nx - 1) * plane + row);
      hz_im = ldv<V>(Hz + (long long)(P.nx - 1) * plane + row);
    } else if (P.x_lo_mode == 2) {
      hy_im = ldv<V>(P.haloH + row);
      hz_im = ldv<V>(P.haloH + plane + row);
    }
  }

  for (int i = ic0; i < ic1; ++i) {
    const long long base = (long long)i * plane;
    Vec<V> hx = zerov<V>(), hy = zerov<V>(), hz = zerov<V>();
    Vec<V> hx_jm = zerov<V>(), hz_jm = zerov<V>();
    Vec<V> ex, ey, ez, ie0, ie1, ie2;
    if (active) {
      hx = ldv<V>(Hx + base + row);
      hy = ldv<V>(Hy + base + row);
      hz = ldv<V>(Hz + base + row);
      if (jm_ok) {
        hx_jm = ldv<V>(Hx + base + rowm);
        hz_jm = ldv<V>(Hz + base + rowm);
      }
      ex = ldv<V>(Ex + base + row);
      ey = ldv<V>(Ey + base + row);
      ez = ldv<V>(Ez + base + row);
      ie0 = ldv<V>(P.eps + base + row);
      if (TIER == 3) {
        ie1 = ldv<V>(P.eps + P.eps_cs + base + row);
        ie2 = ldv<V>(P.eps + 2 * P.eps_cs + base + row);
      } else {
        ie1 = ie0;
        ie2 = ie0;
      }
    }
    // CPML auxiliary fields are fetched together with the field loads (their addresses depend on
    // indices only), so their latency overlaps instead of serialising behind the curl.
    const bool in_x = (PM > 0) && (i < px.lo_len || i >= px.hi_start);
    Vec<V> psx1 = zerov<V>(), psx2 = zerov<V>(), psy1 = zerov<V>(), psy2 = zerov<V>();
    float psz1[V], psz2[V];
    float *qx1 = nullptr, *qx2 = nullptr, *qy1 = nullptr, *qy2 = nullptr;
    if (active) {
      if (in_x) {
        const int side = (i >= px.hi_start) ? 1 : 0;
        const int il = side ? i - px.hi_start : i;
        const long long pidx = ((long long)il * ny + j) * nz + k0;
        qx1 = (side ? px.psiE[1][0] : px.psiE[0][0]) + pidx;
        qx2 = (side ? px.psiE[1][1] : px.psiE[0][1]) + pidx;
        psx1 = ldv<V>(qx1);
        psx2 = ldv<V>(qx2);
      }
      if (in_y) {
        const long long pidx = ((long long)i * yL + jl) * nz + k0;
        qy1 = (yside ? py.psiE[1][0] : py.psiE[0][0]) + pidx;
        qy2 = (yside ? py.psiE[1][1] : py.psiE[0][1]) + pidx;
        psy1 = ldv<V>(qy1);
        psy2 = ldv<V>(qy2);
      }
      if (zvec) {
        if constexpr (V == 4) {
#pragma unroll
          for (int e = 0; e < V; ++e) { psz1[e] = 0.0f; psz2[e] = 0.0f; }
          const float* q1 = pz1 + i * zstride + zoff;
          const float* q2 = pz2 + i * zstride + zoff;
          if (zh0) {
            const float2 t1 = *reinterpret_cast<const float2*>(q1), t2 = *reinterpret_cast<const float2*>(q2);
            psz1[0] = t1.x; psz1[1] = t1.y; psz2[0] = t2.x; psz2[1] = t2.y;
          }
          if (zh1) {
            const float2 t1 = *reinterpret_cast<const float2*>(q1 + 2), t2 = *reinterpret_cast<const float2*>(q2 + 2);
            psz1[2] = t1.x; psz1[3] = t1.y; psz2[2] = t2.x; psz2[3] = t2.y;
          }
        }
      } else if (any_z) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const int k = k0 + e;
          psz1[e] = 0.0f;
          psz2[e] = 0.0f;
          if (k < pz.lo_len || k >= pz.hi_start) {
            const int side = (k >= pz.hi_start) ? 1 : 0;
            const int kl = side ? k - pz.hi_start : k;
            const int L = side ? pz.hi_len : pz.lo_len;
            const long long pidx = ((long long)i * ny + j) * L + kl;
            psz1[e] = (side ? pz.psiE[1][0] : pz.psiE[0][0])[pidx];
            psz2[e] = (side ? pz.psiE[1][1] : pz.psiE[0][1])[pidx];
          }
        }
      }
      // L2 prefetch of this thread's lines FDTDX_PF_DIST planes ahead (holds no registers)
      if (i + FDTDX_PF_DIST < ic1) {
        const long long pb = base + FDTDX_PF_DIST * plane + row;
        prefetch_l2(Hx + pb); prefetch_l2(Hy + pb); prefetch_l2(Hz + pb);
        prefetch_l2(Ex + pb); prefetch_l2(Ey + pb); prefetch_l2(Ez + pb);
        prefetch_l2(P.eps + pb);
        if (TIER == 3) { prefetch_l2(P.eps + P.eps_cs + pb); prefetch_l2(P.eps + 2 * P.eps_cs + pb); }
      }
      if (FDTDX_PF_L1 && i + 1 < ic1) {
        const long long pb = base + plane + row;
        prefetch_l1(Hx + pb); prefetch_l1(Hy + pb); prefetch_l1(Hz + pb);
        prefetch_l1(Ex + pb); prefetch_l1(Ey + pb); prefetch_l1(Ez + pb);
        prefetch_l1(P.eps + pb);
      }
    }
    // z-neighbour (k-1) of the first element: last element of the previous lane
    float hx_l = __shfl_up_sync(0xffffffffu, hx.v[V - 1], 1);
    float hy_l = __shfl_up_sync(0xffffffffu, hy.v[V - 1], 1);
    if (lane == 0) {
      hx_l = (active && km_ok) ? Hx[base + rowk] : 0.0f;
      hy_l = (active && km_ok) ? Hy[base + rowk] : 0.0f;
    }
    if (active) {
      float sBx = 1.0f;
      if (MET) sBx = P.sB[0][i];
      Vec<V> Kx, Ky, Kz;
      Vec<V> dxHz_v, dxHy_v, dyHx_v, dyHz_v, dzHy_v, dzHx_v;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float hx_km = (e == 0) ? hx_l : hx.v[e == 0 ? 0 : e - 1];
        float hy_km = (e == 0) ? hy_l : hy.v[e == 0 ? 0 : e - 1];
        const float sBz = sBzv[e];
        float dyHz = hz.v[e] - hz_jm.v[e];
        float dzHy = hy.v[e] - hy_km;
        float dzHx = hx.v[e] - hx_km;
        float dxHz = hz.v[e] - hz_im.v[e];
        float dxHy = hy.v[e] - hy_im.v[e];
        float dyHx = hx.v[e] - hx_jm.v[e];
        if (MET) {
          dyHz *= sBy; dzHy *= sBz; dzHx *= sBz; dxHz *= sBx; dxHy *= sBx; dyHx *= sBy;
        }
        Kx.v[e] = dyHz - dzHy;
        Ky.v[e] = dzHx - dxHz;
        Kz.v[e] = dxHy - dyHx;
        dxHz_v.v[e] = dxHz; dxHy_v.v[e] = dxHy; dyHx_v.v[e] = dyHx;
        dyHz_v.v[e] = dyHz; dzHy_v.v[e] = dzHy; dzHx_v.v[e] = dzHx;
      }
      // CPML corrections in the reference's object order: x slabs, y slabs, z slabs.
      if (in_x) {
        const float a = px.aE[i], b = px.bE[i], km1 = px.kE[i];
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float c1, c2;  // axis 0: d1 = dx F_z, d2 = dx F_y; corrects K_y (-) and K_z (+)
          cpml_cell(a, b, km1, px.kappa_one, P.simulate && !REV, dxHz_v.v[e], dxHy_v.v[e], &psx1.v[e], &psx2.v[e], &c1, &c2);
          Ky.v[e] = Ky.v[e] - c1;
          Kz.v[e] = Kz.v[e] + c2;
        }
        if (P.simulate && !REV && P.psi_store) { stv<V>(qx1, psx1); stv<V>(qx2, psx2); }
      }
      if (in_y) {
        const float a = py.aE[j], b = py.bE[j], km1 = py.kE[j];
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float c1, c2;  // axis 1: d1 = dy F_x, d2 = dy F_z; corrects K_z (-) and K_x (+)
          cpml_cell(a, b, km1, py.kappa_one, P.simulate && !REV, dyHx_v.v[e], dyHz_v.v[e], &psy1.v[e], &psy2.v[e], &c1, &c2);
          Kz.v[e] = Kz.v[e] - c1;
          Kx.v[e] = Kx.v[e] + c2;
        }
        if (P.simulate && !REV && P.psi_store) { stv<V>(qy1, psy1); stv<V>(qy2, psy2); }
      }
      if (zvec) {
        if constexpr (V == 4) {
#pragma unroll
          for (int e = 0; e < V; ++e) {
            if ((e < 2) ? zh0 : zh1) {
              float c1, c2;  // axis 2: d1 = dz F_y, d2 = dz F_x; corrects K_x (-) and K_y (+)
              cpml_cell(az.v[e], bz.v[e], kz.v[e], pz.kappa_one, P.simulate && !REV, dzHy_v.v[e], dzHx_v.v[e], &psz1[e], &psz2[e], &c1, &c2);
              Kx.v[e] = Kx.v[e] - c1;
              Ky.v[e] = Ky.v[e] + c2;
            }
          }
          if (P.simulate && !REV && P.psi_store) {
            float* q1 = pz1 + i * zstride + zoff;
            float* q2 = pz2 + i * zstride + zoff;
            if (zh0) {
              *reinterpret_cast<float2*>(q1) = make_float2(psz1[0], psz1[1]);
              *reinterpret_cast<float2*>(q2) = make_float2(psz2[0], psz2[1]);
            }
            if (zh1) {
              *reinterpret_cast<float2*>(q1 + 2) = make_float2(psz1[2], psz1[3]);
              *reinterpret_cast<float2*>(q2 + 2) = make_float2(psz2[2], psz2[3]);
            }
          }
        }
      } else if (any_z) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const int k = k0 + e;
          if (k < pz.lo_len || k >= pz.hi_start) {
            const int side = (k >= pz.hi_start) ? 1 : 0;
            const int kl = side ? k - pz.hi_start : k;
            const int L = side ? pz.hi_len : pz.lo_len;
            const long long pidx = ((long long)i * ny + j) * L + kl;
            float c1, c2;  // axis 2: d1 = dz F_y, d2 = dz F_x; corrects K_x (-) and K_y (+)
            cpml_cell(pz.aE[k], pz.bE[k], pz.kE[k], pz.kappa_one, P.simulate && !REV, dzHy_v.v[e], dzHx_v.v[e],
                      &psz1[e], &psz2[e], &c1, &c2);
            if (P.simulate && !REV && P.psi_store) {
              (side ? pz.psiE[1][0] : pz.psiE[0][0])[pidx] = psz1[e];
              (side ? pz.psiE[1][1] : pz.psiE[0][1])[pidx] = psz2[e];
            }
            Kx.v[e] = Kx.v[e] - c1;
            Ky.v[e] = Ky.v[e] + c2;
          }
        }
      }
      // material update
      const bool src_hit = (P.n_src > 0) && any_src_hits(P.src, P.n_src, i, j, k0, V);
      Vec<V> oe[3];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float Eo[3] = {ex.v[e], ey.v[e], ez.v[e]};
        const float K[3] = {Kx.v[e], Ky.v[e], Kz.v[e]};
        const float ie[3] = {ie0.v[e], ie1.v[e], ie2.v[e]};
        const long long cell = base + row + e;
        float En[3];
        if (REV) {
          // update_E_reverse: sources first (inverse), then ((1+s)E - c K inv_eps) / (1-s)
          if (src_hit) inject_E(P.src, P.n_src, P.dt, t, true, i, j, k0 + e, ie[0], ie[1], ie[2], &Eo[0], &Eo[1], &Eo[2]);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float Ec = Eo[c];
            if (SIG) {
              float sg = P.sigE[c * P.sigE_cs + cell];
              float s = (((P.cour * sg) * P.eta0) * ie[c]) / 2.0f;
              Ec = Ec * (1.0f + s);
              En[c] = (Ec - (P.cour * K[c]) * ie[c]) / (1.0f - s);
            } else {
              En[c] = Ec - (P.cour * K[c]) * ie[c];
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float s = 0.0f;
            float E1;
            if (SIG) {
              float sg = P.sigE[c * P.sigE_cs + cell];
              s = (((P.cour * sg) * P.eta0) * ie[c]) / 2.0f;
              E1 = (1.0f - s) * Eo[c] + (P.cour * K[c]) * ie[c];
            } else {
              E1 = Eo[c] + (P.cour * K[c]) * ie[c];
            }
            if (ADE) {
              // P_hat = c1 P + c2 P_prev + c3 E ; E += inv_eps * sum_p (P - P_hat)   (update.py:330-332)
              const long long pstride = 3 * N;
              float delta = 0.0f, c4sum = 0.0f;
              for (int p = 0; p < P.n_poles; ++p) {
                const long long pi = p * pstride + c * N + cell;
                const long long ci = (long long)p * (P.c_cs ? 3 * N : N) + c * P.c_cs + cell;
                float Pc = P.P_cur[pi], Pp = P.P_new[pi];
                float Phat = (P.c1[ci] * Pc + P.c2[ci] * Pp) + P.c3[ci] * Eo[c];
                float dd = Pc - Phat;
                delta = (p == 0) ? dd : delta + dd;
                if (P.has_c4) c4sum = (p == 0) ? P.c4[ci] : c4sum + P.c4[ci];
                P.P_new[pi] = Phat;
              }
              E1 = E1 + ie[c] * delta;
              if (P.has_c4) {
                float divisor = 1.0f + ie[c] * c4sum;
                if (SIG) divisor = divisor + s;
                E1 = E1 / divisor;
                for (int p = 0; p < P.n_poles; ++p) {
                  const long long pi = p * pstride + c * N + cell;
                  const long long ci = (long long)p * (P.c_cs ? 3 * N : N) + c * P.c_cs + cell;
                  P.P_new[pi] = P.P_new[pi] + P.c4[ci] * E1;
                }
              } else if (SIG) {
                E1 = E1 / (1.0f + s);
              }
            } else if (SIG) {
              E1 = E1 / (1.0f + s);
            }
            En[c] = E1;
          }
          if (src_hit) inject_E(P.src, P.n_src, P.dt, t, false, i, j, k0 + e, ie[0], ie[1], ie[2], &En[0], &En[1], &En[2]);
        }
        // PEC walls (pec.py:70-77)
        for (int w = 0; w < P.n_walls; ++w) {
          const WallDev W = P.walls[w];
          if (W.kind == 0 && in_box(W.lo, W.hi, i, j, k0 + e)) {
            if (W.axis != 0) En[0] = 0.0f;
            if (W.axis != 1) En[1] = 0.0f;
            if (W.axis != 2) En[2] = 0.0f;
          }
        }
        oe[0].v[e] = En[0]; oe[1].v[e] = En[1]; oe[2].v[e] = En[2];
      }
      stv<V>(Ex + base + row, oe[0]);
      stv<V>(Ey + base + row, oe[1]);
      stv<V>(Ez + base + row, oe[2]);
    }
    hy_im = hy;
    hz_im = hz;
  }
}

#endif  // !FDTDX_BUILD_H

#if !defined(FDTDX_BUILD_E)
// ------------------------------------------------------------------------------------------------
// H half-step
// ------------------------------------------------------------------------------------------------
template <int V, int MUT, bool REV, bool SIG, bool MET, int PM>
__global__ void __launch_bounds__(256) yee_H_kernel(const StepParams P, const int t) {
  const int lane = threadIdx.x;
  const int k0 = (blockIdx.x * 32 + lane) * V;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int ic0 = P.x_begin + blockIdx.z * P.xchunk;
  const int ic1 = min(ic0 + P.xchunk, P.x_end);
  const bool active = (k0 < P.nz) && (j < P.ny);
  const int nz = P.nz, ny = P.ny;
  const long long plane = (long long)ny * nz;
  const long long N = plane * P.nx;
  const float* __restrict__ Ex = P.E;
  const float* __restrict__ Ey = P.E + N;
  const float* __restrict__ Ez = P.E + 2 * N;
  float* __restrict__ Hx = P.H;
  float* __restrict__ Hy = P.H + N;
  float* __restrict__ Hz = P.H + 2 * N;

  const long long row = (long long)j * nz + k0;
  int jp = j + 1;
  bool jp_ok = true;
  if (jp >= ny) { if (P.wrap[1]) jp = 0; else jp_ok = false; }
  const long long rowp = (long long)jp * nz + k0;
  int kp = k0 + V;  // first cell after this thread's vector
  bool kp_ok = true;
  if (kp >= nz) { if (P.wrap[2]) kp = 0; else kp_ok = false; }
  const long long rowk = (long long)j * nz + kp;
  const bool last_lane = (lane == 31) || (k0 + V >= nz);

  const AxisPmlDev& px = P.pml[0];
  const AxisPmlDev& py = P.pml[1];
  const AxisPmlDev& pz = P.pml[2];
  // PM: 0 = no CPML slab on this plan (all CPML code compiled out), 1 = scalar z-slab path, 2 = 128-bit z-slab path
  const bool in_y = (PM > 0) && active && (j < py.lo_len || j >= py.hi_start);
  const int yside = (j >= py.hi_start) ? 1 : 0;
  const int jl = yside ? j - py.hi_start : j;
  const int yL = yside ? py.hi_len : py.lo_len;
  const bool any_z = (PM > 0) && active && (k0 < pz.lo_len || k0 + V > pz.hi_start);
  // PM == 2 (V == 4, even slab thickness): the lane's four cells are two 64-bit halves, each either
  // entirely inside or entirely outside a z slab, so psi moves as float2 with no per-cell branch.
  const int zside = (k0 + V > pz.hi_start) ? 1 : 0;
  const bool zvec = (PM == 2) && any_z;
  const bool zh0 = zvec && (zside ? (k0 >= pz.hi_start) : (k0 < pz.lo_len));
  const bool zh1 = zvec && (zside ? (k0 + 2 >= pz.hi_start) : (k0 + 2 < pz.lo_len));
  const long long zstride = (long long)ny * (zside ? pz.hi_len : pz.lo_len);
  const long long zoff = (long long)j * (zside ? pz.hi_len : pz.lo_len) + (zside ? k0 - pz.hi_start : k0);
  float* const pz1 = zside ? pz.psiH[1][0] : pz.psiH[0][0];
  float* const pz2 = zside ? pz.psiH[1][1] : pz.psiH[0][1];
  Vec<V> az = zerov<V>(), bz = zerov<V>(), kz = zerov<V>();
  if (zvec) { az = ldv<V>(pz.aH + k0); bz = ldv<V>(pz.bH + k0); kz = ldv<V>(pz.kH + k0); }
  float sFy = 1.0f;
  float sFzv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) sFzv[e] = 1.0f;
  if (MET && active) {
    sFy = P.sF[1][j];
#pragma unroll
    for (int e = 0; e < V; ++e) sFzv[e] = P.sF[2][k0 + e];
  }

  // register queue: E of the current plane is the "next" plane loaded one step earlier
  Vec<V> ex = zerov<V>(), ey = zerov<V>(), ez = zerov<V>();
  if (active) {
    ex = ldv<V>(Ex + (long long)ic0 * plane + row);
    ey = ldv<V>(Ey + (long long)ic0 * plane + row);
    ez = ldv<V>(Ez + (long long)ic0 * plane + row);
  }

  for (int i = ic0; i < ic1; ++i) {
    const long long base = (long long)i * plane;
    Vec<V> ex_n = zerov<V>(), ey_n = zerov<V>(), ez_n = zerov<V>();
    Vec<V> ex_jp = zerov<V>(), ez_jp = zerov<V>();
    Vec<V> hx, hy, hz, im0, im1, im2;
    if (active) {
      if (i + 1 < P.nx) {
        const long long bn = base + plane;
        ey_n = ldv<V>(Ey + bn + row);
        ez_n = ldv<V>(Ez + bn + row);
        if (i + 1 < ic1) ex_n = ldv<V>(Ex + bn + row);
      } else if (P.x_hi_mode == 1) {
        ey_n = ldv<V>(Ey + row);
        ez_n = ldv<V>(Ez + row);
      } else if (P.x_hi_mode == 2) {
        ey_n = ldv<V>(P.haloE + row);
        ez_n = ldv<V>(P.haloE + plane + row);
      }
      if (jp_ok) {
        ex_jp = ldv<V>(Ex + base + rowp);
        ez_jp = ldv<V>(Ez + base + rowp);
      }
      hx = ldv<V>(Hx + base + row);
      hy = ldv<V>(Hy + base + row);
      hz = ldv<V>(Hz + base + row);
      if (MUT >= 1) {
        im0 = ldv<V>(P.mu + base + row);
        if (MUT == 3) {
          im1 = ldv<V>(P.mu + P.mu_cs + base + row);
          im2 = ldv<V>(P.mu + 2 * P.mu_cs + base + row);
        } else {
          im1 = im0;
          im2 = im0;
        }
      }
    }
    // CPML auxiliary fields are fetched together with the field loads (their addresses depend on
    // indices only), so their latency overlaps instead of serialising behind the curl.
    const bool in_x = (PM > 0) && (i < px.lo_len || i >= px.hi_start);
    Vec<V> psx1 = zerov<V>(), psx2 = zerov<V>(), psy1 = zerov<V>(), psy2 = zerov<V>();
    float psz1[V], psz2[V];
    float *qx1 = nullptr, *qx2 = nullptr, *qy1 = nullptr, *qy2 = nullptr;
    if (active) {
      if (in_x) {
        const int side = (i >= px.hi_start) ? 1 : 0;
        const int il = side ? i - px.hi_start : i;
        const long long pidx = ((long long)il * ny + j) * nz + k0;
        qx1 = (side ? px.psiH[1][0] : px.psiH[0][0]) + pidx;
        qx2 = (side ? px.psiH[1][1] : px.psiH[0][1]) + pidx;
        psx1 = ldv<V>(qx1);
        psx2 = ldv<V>(qx2);
      }
      if (in_y) {
        const long long pidx = ((long long)i * yL + jl) * nz + k0;
        qy1 = (yside ? py.psiH[1][0] : py.psiH[0][0]) + pidx;
        qy2 = (yside ? py.psiH[1][1] : py.psiH[0][1]) + pidx;
        psy1 = ldv<V>(qy1);
        psy2 = ldv<V>(qy2);
      }
      if (zvec) {
        if constexpr (V == 4) {
#pragma unroll
          for (int e = 0; e < V; ++e) { psz1[e] = 0.0f; psz2[e] = 0.0f; }
          const float* q1 = pz1 + i * zstride + zoff;
          const float* q2 = pz2 + i * zstride + zoff;
          if (zh0) {
            const float2 t1 = *reinterpret_cast<const float2*>(q1), t2 = *reinterpret_cast<const float2*>(q2);
            psz1[0] = t1.x; psz1[1] = t1.y; psz2[0] = t2.x; psz2[1] = t2.y;
          }
          if (zh1) {
            const float2 t1 = *reinterpret_cast<const float2*>(q1 + 2), t2 = *reinterpret_cast<const float2*>(q2 + 2);
            psz1[2] = t1.x; psz1[3] = t1.y; psz2[2] = t2.x; psz2[3] = t2.y;
          }
        }
      } else if (any_z) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const int k = k0 + e;
          psz1[e] = 0.0f;
          psz2[e] = 0.0f;
          if (k < pz.lo_len || k >= pz.hi_start) {
            const int side = (k >= pz.hi_start) ? 1 : 0;
            const int kl = side ? k - pz.hi_start : k;
            const int L = side ? pz.hi_len : pz.lo_len;
            const long long pidx = ((long long)i * ny + j) * L + kl;
            psz1[e] = (side ? pz.psiH[1][0] : pz.psiH[0][0])[pidx];
            psz2[e] = (side ? pz.psiH[1][1] : pz.psiH[0][1])[pidx];
          }
        }
      }
      // L2 prefetch of this thread's lines FDTDX_PF_DIST planes ahead (holds no registers)
      if (i + FDTDX_PF_DIST < ic1) {
        const long long pb = base + FDTDX_PF_DIST * plane + row;
        prefetch_l2(Hx + pb); prefetch_l2(Hy + pb); prefetch_l2(Hz + pb);
        prefetch_l2(Ex + pb); prefetch_l2(Ey + pb); prefetch_l2(Ez + pb);
        if (MUT >= 1) prefetch_l2(P.mu + pb);
        if (MUT == 3) { prefetch_l2(P.mu + P.mu_cs + pb); prefetch_l2(P.mu + 2 * P.mu_cs + pb); }
      }
      if (FDTDX_PF_L1 && i + 2 < ic1) {
        const long long pb = base + 2 * plane + row;
        prefetch_l1(Ex + pb); prefetch_l1(Ey + pb); prefetch_l1(Ez + pb);
        const long long ph = base + plane + row;
        prefetch_l1(Hx + ph); prefetch_l1(Hy + ph); prefetch_l1(Hz + ph);
      }
    }
    float ex_r = __shfl_down_sync(0xffffffffu, ex.v[0], 1);
    float ey_r = __shfl_down_sync(0xffffffffu, ey.v[0], 1);
    if (last_lane) {
      ex_r = (active && kp_ok) ? Ex[base + rowk] : 0.0f;
      ey_r = (active && kp_ok) ? Ey[base + rowk] : 0.0f;
    }
    if (active) {
      float sFx = 1.0f;
      if (MET) sFx = P.sF[0][i];
      Vec<V> Kx, Ky, Kz;
      Vec<V> dxEz_v, dxEy_v, dyEx_v, dyEz_v, dzEy_v, dzEx_v;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float ex_kp = (e == V - 1) ? ex_r : ex.v[e == V - 1 ? e : e + 1];
        float ey_kp = (e == V - 1) ? ey_r : ey.v[e == V - 1 ? e : e + 1];
        const float sFz = sFzv[e];
        float dyEz = ez_jp.v[e] - ez.v[e];
        float dzEy = ey_kp - ey.v[e];
        float dzEx = ex_kp - ex.v[e];
        float dxEz = ez_n.v[e] - ez.v[e];
        float dxEy = ey_n.v[e] - ey.v[e];
        float dyEx = ex_jp.v[e] - ex.v[e];
        if (MET) {
          dyEz *= sFy; dzEy *= sFz; dzEx *= sFz; dxEz *= sFx; dxEy *= sFx; dyEx *= sFy;
        }
        Kx.v[e] = dyEz - dzEy;
        Ky.v[e] = dzEx - dxEz;
        Kz.v[e] = dxEy - dyEx;
        dxEz_v.v[e] = dxEz; dxEy_v.v[e] = dxEy; dyEx_v.v[e] = dyEx;
        dyEz_v.v[e] = dyEz; dzEy_v.v[e] = dzEy; dzEx_v.v[e] = dzEx;
      }
      if (in_x) {
        const float a = px.aH[i], b = px.bH[i], km1 = px.kH[i];
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float c1, c2;  // axis 0: d1 = dx F_z, d2 = dx F_y; corrects K_y (-) and K_z (+)
          cpml_cell(a, b, km1, px.kappa_one, P.simulate && !REV, dxEz_v.v[e], dxEy_v.v[e], &psx1.v[e], &psx2.v[e], &c1, &c2);
          Ky.v[e] = Ky.v[e] - c1;
          Kz.v[e] = Kz.v[e] + c2;
        }
        if (P.simulate && !REV && P.psi_store) { stv<V>(qx1, psx1); stv<V>(qx2, psx2); }
      }
      if (in_y) {
        const float a = py.aH[j], b = py.bH[j], km1 = py.kH[j];
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float c1, c2;  // axis 1: d1 = dy F_x, d2 = dy F_z; corrects K_z (-) and K_x (+)
          cpml_cell(a, b, km1, py.kappa_one, P.simulate && !REV, dyEx_v.v[e], dyEz_v.v[e], &psy1.v[e], &psy2.v[e], &c1, &c2);
          Kz.v[e] = Kz.v[e] - c1;
          Kx.v[e] = Kx.v[e] + c2;
        }
        if (P.simulate && !REV && P.psi_store) { stv<V>(qy1, psy1); stv<V>(qy2, psy2); }
      }
      if (zvec) {
        if constexpr (V == 4) {
#pragma unroll
          for (int e = 0; e < V; ++e) {
            if ((e < 2) ? zh0 : zh1) {
              float c1, c2;  // axis 2: d1 = dz F_y, d2 = dz F_x; corrects K_x (-) and K_y (+)
              cpml_cell(az.v[e], bz.v[e], kz.v[e], pz.kappa_one, P.simulate && !REV, dzEy_v.v[e], dzEx_v.v[e], &psz1[e], &psz2[e], &c1, &c2);
              Kx.v[e] = Kx.v[e] - c1;
              Ky.v[e] = Ky.v[e] + c2;
            }
          }
          if (P.simulate && !REV && P.psi_store) {
            float* q1 = pz1 + i * zstride + zoff;
            float* q2 = pz2 + i * zstride + zoff;
            if (zh0) {
              *reinterpret_cast<float2*>(q1) = make_float2(psz1[0], psz1[1]);
              *reinterpret_cast<float2*>(q2) = make_float2(psz2[0], psz2[1]);
            }
            if (zh1) {
              *reinterpret_cast<float2*>(q1 + 2) = make_float2(psz1[2], psz1[3]);
              *reinterpret_cast<float2*>(q2 + 2) = make_float2(psz2[2], psz2[3]);
            }
          }
        }
      } else if (any_z) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const int k = k0 + e;
          if (k < pz.lo_len || k >= pz.hi_start) {
            const int side = (k >= pz.hi_start) ? 1 : 0;
            const int kl = side ? k - pz.hi_start : k;
            const int L = side ? pz.hi_len : pz.lo_len;
            const long long pidx = ((long long)i * ny + j) * L + kl;
            float c1, c2;  // axis 2: d1 = dz F_y, d2 = dz F_x; corrects K_x (-) and K_y (+)
            cpml_cell(pz.aH[k], pz.bH[k], pz.kH[k], pz.kappa_one, P.simulate && !REV, dzEy_v.v[e], dzEx_v.v[e],
                      &psz1[e], &psz2[e], &c1, &c2);
            if (P.simulate && !REV && P.psi_store) {
              (side ? pz.psiH[1][0] : pz.psiH[0][0])[pidx] = psz1[e];
              (side ? pz.psiH[1][1] : pz.psiH[0][1])[pidx] = psz2[e];
            }
            Kx.v[e] = Kx.v[e] - c1;
            Ky.v[e] = Ky.v[e] + c2;
          }
        }
      }
      const bool src_hit = (P.n_src > 0) && any_src_hits(P.src, P.n_src, i, j, k0, V);
      Vec<V> oh[3];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float Ho[3] = {hx.v[e], hy.v[e], hz.v[e]};
        const float K[3] = {Kx.v[e], Ky.v[e], Kz.v[e]};
        float im[3];
        if (MUT >= 1) { im[0] = im0.v[e]; im[1] = im1.v[e]; im[2] = im2.v[e]; }
        else { im[0] = im[1] = im[2] = P.inv_mu_scalar; }
        const long long cell = base + row + e;
        float Hn[3];
        if (REV) {
          if (src_hit) inject_H(P.src, P.n_src, P.dt, t, true, i, j, k0 + e, im[0], im[1], im[2], &Ho[0], &Ho[1], &Ho[2]);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float Hc = Ho[c];
            if (SIG) {
              float sg = P.sigH[c * P.sigH_cs + cell];
              float s = (((P.cour * sg) / P.eta0) * im[c]) / 2.0f;
              Hc = Hc * (1.0f + s);
              Hn[c] = (Hc + (P.cour * K[c]) * im[c]) / (1.0f - s);
            } else {
              Hn[c] = Hc + (P.cour * K[c]) * im[c];
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (SIG) {
              float sg = P.sigH[c * P.sigH_cs + cell];
              float s = (((P.cour * sg) / P.eta0) * im[c]) / 2.0f;
              float H1 = (1.0f - s) * Ho[c] - (P.cour * K[c]) * im[c];
              Hn[c] = H1 / (1.0f + s);
            } else {
              Hn[c] = Ho[c] - (P.cour * K[c]) * im[c];
            }
          }
          if (src_hit) inject_H(P.src, P.n_src, P.dt, t, false, i, j, k0 + e, im[0], im[1], im[2], &Hn[0], &Hn[1], &Hn[2]);
        }
        for (int w = 0; w < P.n_walls; ++w) {
          const WallDev W = P.walls[w];
          if (W.kind == 1 && in_box(W.lo, W.hi, i, j, k0 + e)) {
            if (W.axis != 0) Hn[0] = 0.0f;
            if (W.axis != 1) Hn[1] = 0.0f;
            if (W.axis != 2) Hn[2] = 0.0f;
          }
        }
        oh[0].v[e] = Hn[0]; oh[1].v[e] = Hn[1]; oh[2].v[e] = Hn[2];
      }
      stv<V>(Hx + base + row, oh[0]);
      stv<V>(Hy + base + row, oh[1]);
      stv<V>(Hz + base + row, oh[2]);
    }
    ex = ex_n;
    ey = ey_n;
    ez = ez_n;
  }
}
#endif  // !FDTDX_BUILD_E
