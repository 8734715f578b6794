// Detector accumulation for LARGE exact-interpolation regions (DET_VOLUME): full-volume energy videos,
// field videos, volume reductions (update_detector_states, fdtd/update.py:1040-1137; interpolate_fields,
// core/physics/curl.py:86-224; Detector.update of objects/detectors/*.py).
//
// The generic kernels of aux_kernels.cuh spend ~900 instructions per cell (one thread per cell, 64-bit
// index arithmetic, 30 scalar loads, a scalar H_prev copy).  Here a warp owns one (x, y) row of the
// region and a lane four consecutive z cells:
//   det_gather_rows_kernel  copies the pre-update H rows (halo rules applied) into a 16-byte aligned box,
//   det_march_kernel        reads every operand row of the co-location stencil with one 128-bit load
//                           (k+1 by shuffle), evaluates the same expressions in the same order as
//                           colocate_interior / colocate_t (bit-identical samples), and hands each cell to
//                           det_emit - or, for averaged energy slices, reduces the three means itself
//                           (row sum by warp shuffle, y sum over the CTA's rows, x sum along the march)
//                           into small partial buffers,
//   det_mean_finish_kernel  folds those partials in a fixed order and writes the three planes.
// Sums are deterministic (fixed geometry, fixed order); they differ from the sequential order of
// det_slice_mean_kernel by float32 rounding only (~1e-7 relative).
#pragma once
#include "aux_kernels.cuh"

#define DETV_ROWS 8     // y rows (warps) per CTA
#define DETV_TZ 128     // z cells per warp pass
#define DETV_XC 4       // x planes per CTA

// aligned H_prev box: element (c, a, b, g) holds H_c(lo_x - 1 + a, lo_y - 1 + b, hz0 + g)
__device__ __forceinline__ long long detv_hidx(const DetDev& D, int c, int a, int b) {
  const int sy = D.hi[1] - D.lo[1] + 1, sx = D.hi[0] - D.lo[0] + 1;
  return (((long long)c * sx + a) * sy + b) * D.hrow;
}

// One warp per (c, a, b) row.  Source rows outside the grid follow the halo rule (zero / wrap).
__global__ void __launch_bounds__(256) det_gather_rows_kernel(const GridDev G, const DetDev* __restrict__ dets, const int t, const int inverse) {
  __shared__ DetDev sD;
  det_stage_descriptor(&sD, dets + blockIdx.y);
  const DetDev& D = sD;
  if (((D.flags & DET_INVERSE) != 0) != (inverse != 0) || !D.on[t] || !(D.flags & DET_VOLUME)) return;
  const int sx = D.hi[0] - D.lo[0] + 1, sy = D.hi[1] - D.lo[1] + 1;
  const long long rows = 3LL * sx * sy;
  const int lane = threadIdx.x & 31;
  const long long N = (long long)G.nx * G.ny * G.nz;
  for (long long r = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (blockDim.x >> 5)) {  // grid covers all rows: one row per warp
    const int b = (int)(r % sy);
    const int a = (int)((r / sy) % sx);
    const int c = (int)(r / ((long long)sy * sx));
    int x = D.lo[0] - 1 + a, y = D.lo[1] - 1 + b;
    bool zero = false;
    if (x < 0) { if (G.wrap[0]) x += G.nx; else zero = true; }
    if (y < 0) { if (G.wrap[1]) y += G.ny; else zero = true; }
    float4* dst = reinterpret_cast<float4*>(D.hprev + detv_hidx(D, c, a, b));
    const float* src = G.H + c * N + ((long long)x * G.ny + y) * G.nz;
    for (int q = lane; q < D.hrow / 4; q += 32) {
      const int z = D.hz0 + 4 * q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!zero) {
        if (z + 3 < G.nz) {
          v = *reinterpret_cast<const float4*>(src + z);
        } else {  // the row's tail: z+1 of the last cell is the halo (zero / wrap to z = 0)
          float e[4];
          for (int k = 0; k < 4; ++k) e[k] = (z + k < G.nz) ? src[z + k] : ((G.wrap[2] && z + k < G.nz + 4) ? src[z + k - G.nz] : 0.0f);
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      dst[q] = v;
    }
  }
}

struct Row4 {
  float v[4];
  float nx;  // the element after v[3] (k+1 of the lane's last cell)
};
// A row operand is fetched in two phases so that the ~20 row loads of a plane are all in flight before
// the first warp shuffle (a shuffle right after each load serialises the round trips: ncu showed 12 us
// per plane).  Phase 1 (RowLd): the 128-bit load plus, for the lane that cannot get k+1 from its
// neighbour, the scalar behind it.  Phase 2 (row_finish): k+1 from the next lane.
struct RowLd {
  float4 t;
  float tail;
};
__device__ __forceinline__ Row4 row_finish(const RowLd& l, const bool want_next, const bool own_tail) {
  Row4 r;
  r.v[0] = l.t.x; r.v[1] = l.t.y; r.v[2] = l.t.z; r.v[3] = l.t.w;
  r.nx = 0.0f;
  if (want_next) {
    const float nxt = __shfl_down_sync(0xffffffffu, l.t.x, 1);
    r.nx = own_tail ? l.tail : nxt;
  }
  return r;
}
__device__ __forceinline__ float row_at(const Row4& r, int e) { return e < 4 ? r.v[e] : r.nx; }
// H_bar = (H_prev + H) / 2 per element (update.py:1088, 1098)
__device__ __forceinline__ Row4 detv_hbar(const Row4& p, const Row4& h) {
  Row4 r;
#pragma unroll
  for (int e = 0; e < 4; ++e) r.v[e] = (p.v[e] + h.v[e]) / 2.0f;
  r.nx = (p.nx + h.nx) / 2.0f;
  return r;
}

// Addresses of the four stencil rows (x, y), (x-1, y), (x, y-1), (x-1, y-1) of one thread: element
// offsets into a (Nx, Ny, Nz) component, or "zero row" where the halo rule says so.
struct RowSet {
  long long o[4];
  bool zero[4];
};

// grid: (z tiles of 128, y tiles of 8 rows, x chunks of DETV_XC planes) of the detector box; blockIdx.z
// also enumerates detectors: z = det * nxc + chunk.  No CTA-wide barrier in the plane loop: warps are
// independent (slice means: the z sum is a warp reduction, the y / x sums are folded from the per-cell
// energies by det_mean_finish_kernel).
// MODE 1: energy detectors on isotropic / diagonal media - the energy density is evaluated inline
// (metrics.py:55-67, same expressions as det_emit) with 128-bit material loads and one division per
// distinct material component; MODE 0: every other kind goes through det_emit.
__device__ __forceinline__ int detv_mode(const GridDev& G, const DetDev& D) {
  return (D.kind == 1 && G.eps_tier != 9 && G.mu_tier != 9) ? 1 : 0;
}
template <bool EXACT, int MODE>
__global__ void __launch_bounds__(256, 2) det_march_kernel(const GridDev G, const DetDev* __restrict__ dets, const int t, const int inverse, const int nxc_max) {
  __shared__ DetDev sD;
  const int di = blockIdx.z / nxc_max, xc = blockIdx.z - di * nxc_max;
  {
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int q = tid; q < (int)(sizeof(DetDev) / 4); q += 256) reinterpret_cast<int*>(&sD)[q] = reinterpret_cast<const int*>(dets + di)[q];
    __syncthreads();
  }
  const DetDev& D = sD;
  if (((D.flags & DET_INVERSE) != 0) != (inverse != 0) || !D.on[t] || !(D.flags & DET_VOLUME)) return;
  if (((D.flags & DET_EXACT) != 0) != EXACT || detv_mode(G, D) != MODE) return;
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const int lane = threadIdx.x, wrow = threadIdx.y;
  // lanes are aligned to global multiples of 4 in z; cells outside [lo_z, hi_z) are masked at emit
  const int zt0 = (D.lo[2] & ~3) + blockIdx.x * DETV_TZ;
  const int z0 = zt0 + 4 * lane;
  const int ry = blockIdx.y * DETV_ROWS + wrow;
  const int rx0 = xc * DETV_XC, rx1 = min(rx0 + DETV_XC, ex);
  if (zt0 >= D.hi[2] || rx0 >= ex || ry >= ey) return;  // warp-uniform
  const int y = D.lo[1] + ry;
  const bool fused_mean = (D.flags & DET_SLICES) && (D.flags & DET_SLICE_MEAN) && D.kind == 1;
  const bool z_in = z0 < G.nz;                                        // this lane holds grid cells
  const bool lane_in = z_in && z0 < D.hi[2] && z0 + 4 > D.lo[2];      // ... and cells of the region
  const bool hp_in = (z0 - D.hz0 + 4 <= D.hrow);                      // inside the gathered row (covers hi_z)
  const bool own_tail = (lane == 31) || (z0 + 4 >= G.nz);             // k+1 of the last cell is not in the next lane
  const long long plane = (long long)G.ny * G.nz, N = plane * G.nx;
  // y-neighbour row (halo rule on the min-y face)
  long long dy = -(long long)G.nz;
  bool y_zero = false;
  if (y == 0) { if (G.wrap[1]) dy = (long long)(G.ny - 1) * G.nz; else y_zero = true; }
  // H_prev box strides
  const int hsy = ey + 1;
  const long long h_row = D.hrow, h_pl = (long long)hsy * D.hrow, h_c = (long long)(ex + 1) * h_pl;
  const float* const hp0 = D.hprev + ((long long)(ry + 1)) * h_row + (z0 - D.hz0);  // (c=0, a=0, b=ry+1)

  auto ld = [&](const float* F, const long long off, const bool zero, const bool want_next) {
    RowLd r;
    r.t = make_float4(0.f, 0.f, 0.f, 0.f);
    r.tail = 0.0f;
    if (!zero && z_in) {
      r.t = *reinterpret_cast<const float4*>(F + off);
      if (want_next && own_tail) {
        if (z0 + 4 < G.nz) r.tail = F[off + 4];
        else if (G.wrap[2]) r.tail = F[off - z0];
      }
    }
    return r;
  };
  auto ldp = [&](const float* q, const bool want_next) {
    RowLd r;
    r.t = make_float4(0.f, 0.f, 0.f, 0.f);
    r.tail = 0.0f;
    if (hp_in) {
      r.t = *reinterpret_cast<const float4*>(q);
      if (want_next && lane == 31 && (z0 + 4 - D.hz0) < D.hrow) r.tail = q[4];
    }
    return r;
  };

  for (int rx = rx0; rx < rx1; ++rx) {
    const int x = D.lo[0] + rx;
    const long long o_c = ((long long)x * G.ny + y) * G.nz + z0;
    long long dx = -plane;
    bool x_zero = false;
    if (x == 0) { if (G.wrap[0]) dx = (long long)(G.nx - 1) * plane; else x_zero = true; }
    float Es[4][3], Hs[4][3];
    if (EXACT) {
      const float *E0 = G.E, *E1 = G.E + N, *E2 = G.E + 2 * N, *H0 = G.H, *H1 = G.H + N, *H2 = G.H + 2 * N;
      const float* hp = hp0 + (long long)(rx + 1) * h_pl;  // (c=0, a=rx+1, b=ry+1)
      // phase 1: every load of this plane
      const RowLd l_exc = ld(E0, o_c, false, true), l_exm = ld(E0, o_c + dx, x_zero, true);
      const RowLd l_eyc = ld(E1, o_c, false, true), l_eym = ld(E1, o_c + dy, y_zero, true);
      const RowLd l_ezc = ld(E2, o_c, false, false);
      const RowLd n_hxc = ld(H0, o_c, false, false), n_hxm = ld(H0, o_c + dy, y_zero, false);
      const RowLd n_hyc = ld(H1, o_c, false, false), n_hym = ld(H1, o_c + dx, x_zero, false);
      const RowLd n_zcc = ld(H2, o_c, false, true), n_zmc = ld(H2, o_c + dx, x_zero, true);
      const RowLd n_zcm = ld(H2, o_c + dy, y_zero, true), n_zmm = ld(H2, o_c + dx + dy, x_zero || y_zero, true);
      const RowLd p_hxc = ldp(hp, false), p_hxm = ldp(hp - h_row, false);
      const RowLd p_hyc = ldp(hp + h_c, false), p_hym = ldp(hp + h_c - h_pl, false);
      const RowLd p_zcc = ldp(hp + 2 * h_c, true), p_zmc = ldp(hp + 2 * h_c - h_pl, true);
      const RowLd p_zcm = ldp(hp + 2 * h_c - h_row, true), p_zmm = ldp(hp + 2 * h_c - h_pl - h_row, true);
      // phase 2: k+1 neighbours by shuffle, time-centred H
      const Row4 exc = row_finish(l_exc, true, own_tail), exm = row_finish(l_exm, true, own_tail);
      const Row4 eyc = row_finish(l_eyc, true, own_tail), eym = row_finish(l_eym, true, own_tail);
      const Row4 ezc = row_finish(l_ezc, false, false);
      const Row4 hx_c = detv_hbar(row_finish(p_hxc, false, false), row_finish(n_hxc, false, false));
      const Row4 hx_m = detv_hbar(row_finish(p_hxm, false, false), row_finish(n_hxm, false, false));
      const Row4 hy_c = detv_hbar(row_finish(p_hyc, false, false), row_finish(n_hyc, false, false));
      const Row4 hy_m = detv_hbar(row_finish(p_hym, false, false), row_finish(n_hym, false, false));
      const Row4 hz_cc = detv_hbar(row_finish(p_zcc, true, lane == 31), row_finish(n_zcc, true, own_tail));
      const Row4 hz_mc = detv_hbar(row_finish(p_zmc, true, lane == 31), row_finish(n_zmc, true, own_tail));
      const Row4 hz_cm = detv_hbar(row_finish(p_zcm, true, lane == 31), row_finish(n_zcm, true, own_tail));
      const Row4 hz_mm = detv_hbar(row_finish(p_zmm, true, lane == 31), row_finish(n_zmm, true, own_tail));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // same expressions, same order as colocate_interior / colocate_t (curl.py:120-222)
        float lo = bea(G, exc.v[e], exm.v[e], 0, x);
        float hi = bea(G, row_at(exc, e + 1), row_at(exm, e + 1), 0, x);
        Es[e][0] = (lo + hi) / 2.0f;
        lo = bea(G, eyc.v[e], eym.v[e], 1, y);
        hi = bea(G, row_at(eyc, e + 1), row_at(eym, e + 1), 1, y);
        Es[e][1] = (lo + hi) / 2.0f;
        Es[e][2] = ezc.v[e];
        Hs[e][0] = bea(G, hx_c.v[e], hx_m.v[e], 1, y);
        Hs[e][1] = bea(G, hy_c.v[e], hy_m.v[e], 0, x);
        const float lx = bea(G, hz_cc.v[e], hz_mc.v[e], 0, x);
        const float lxm = bea(G, hz_cm.v[e], hz_mm.v[e], 0, x);
        const float lxy = bea(G, lx, lxm, 1, y);
        const float hx2 = bea(G, row_at(hz_cc, e + 1), row_at(hz_mc, e + 1), 0, x);
        const float hxm = bea(G, row_at(hz_cm, e + 1), row_at(hz_mm, e + 1), 0, x);
        const float hxy = bea(G, hx2, hxm, 1, y);
        Hs[e][2] = (lxy + hxy) / 2.0f;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const RowLd a = ld(G.E + c * N, o_c, false, false), b = ld(G.H + c * N, o_c, false, false);
        Es[0][c] = a.t.x; Es[1][c] = a.t.y; Es[2][c] = a.t.z; Es[3][c] = a.t.w;
        Hs[0][c] = b.t.x; Hs[1][c] = b.t.y; Hs[2][c] = b.t.z; Hs[3][c] = b.t.w;
      }
    }
    float ev[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE == 1) {
      if (lane_in) {
        // 0.5 * (1 / inv_eps_c) and 0.5 * (1 / inv_mu_c) per component (one division per distinct array)
        float he[3][4], hm[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c == 0 || G.eps_tier == 3) {
            const float4 q = *reinterpret_cast<const float4*>(G.eps + c * G.eps_cs + o_c);
            he[c][0] = 0.5f * (1.0f / q.x); he[c][1] = 0.5f * (1.0f / q.y); he[c][2] = 0.5f * (1.0f / q.z); he[c][3] = 0.5f * (1.0f / q.w);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) he[c][e] = he[0][e];
          }
          if (G.mu == nullptr) {
            const float h = 0.5f * (1.0f / G.inv_mu_scalar);
#pragma unroll
            for (int e = 0; e < 4; ++e) hm[c][e] = h;
          } else if (c == 0 || G.mu_tier == 3) {
            const float4 q = *reinterpret_cast<const float4*>(G.mu + c * G.mu_cs + o_c);
            hm[c][0] = 0.5f * (1.0f / q.x); hm[c][1] = 0.5f * (1.0f / q.y); hm[c][2] = 0.5f * (1.0f / q.z); hm[c][3] = 0.5f * (1.0f / q.w);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) hm[c][e] = hm[0][e];
          }
        }
        const bool staged = (D.flags & DET_REDUCE) != 0;
        const int slot = D.arr_idx[t];
        const long long n_cells = (long long)ex * ey * ez;
        const long long cell0 = ((long long)rx * ey + ry) * ez + (z0 - D.lo[2]);
        float* const dst = (fused_mean || staged) ? D.scratch : D.state[0] + (long long)slot * n_cells;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int z = z0 + e;
          if (z < D.lo[2] || z >= D.hi[2]) continue;
          // eE = sum_c 0.5/inv_eps_c * |E_c|^2 accumulated as (x + y) + z, likewise eH (metrics.py:58-66)
          const float a0 = he[0][e] * (fabsf(Es[e][0]) * fabsf(Es[e][0])), a1 = he[1][e] * (fabsf(Es[e][1]) * fabsf(Es[e][1]));
          const float a2 = he[2][e] * (fabsf(Es[e][2]) * fabsf(Es[e][2]));
          const float b0 = hm[0][e] * (fabsf(Hs[e][0]) * fabsf(Hs[e][0])), b1 = hm[1][e] * (fabsf(Hs[e][1]) * fabsf(Hs[e][1]));
          const float b2 = hm[2][e] * (fabsf(Hs[e][2]) * fabsf(Hs[e][2]));
          const float en = ((a0 + a1) + a2) + ((b0 + b1) + b2);
          ev[e] = en;
          dst[cell0 + e] = en;
        }
      }
    } else if (lane_in) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int z = z0 + e;
        if (z < D.lo[2] || z >= D.hi[2]) continue;
        const int rz = z - D.lo[2];
        const long long cell = ((long long)rx * ey + ry) * ez + rz;
        det_emit<false>(G, D, t, cell, rx, ry, rz, x, y, z, Es[e], Hs[e], nullptr);
      }
    }
    if (fused_mean) {  // warp-uniform: sum over z of this row's 128-cell pass -> part[0][ztile][rx][ry]
      float rs = (ev[0] + ev[1]) + (ev[2] + ev[3]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      if (lane == 0) D.part[0][((long long)blockIdx.x * ex + rx) * ey + ry] = rs;
    }
  }
}

// Averaged energy slices: the XY plane folds the per-row z sums, the XZ / YZ planes sum the staged
// per-cell energies over y / x (consecutive threads read consecutive z: coalesced), each in a fixed order.
__global__ void det_mean_finish_kernel(const DetDev* __restrict__ dets, const int di, const int t) {
  const DetDev D = dets[di];
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const long long nxy = (long long)ex * ey, nxz = (long long)ex * ez, nyz = (long long)ey * ez;
  const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int slot = D.arr_idx[t];
  if (o < nxz) {  // XZ plane: mean over y
    const int rx = (int)(o / ez), rz = (int)(o - (long long)rx * ez);
    const float* p = D.scratch + (long long)rx * ey * ez + rz;
    float a = 0.0f;
#pragma unroll 8
    for (int q = 0; q < ey; ++q) a += p[(long long)q * ez];
    D.state[1][slot * nxz + o] = a / (float)ey;
  } else if (o < nxz + nyz) {  // YZ plane: mean over x
    const long long i = o - nxz;
    const float* p = D.scratch + i;
    float a = 0.0f;
#pragma unroll 8
    for (int q = 0; q < ex; ++q) a += p[(long long)q * nyz];
    D.state[2][slot * nyz + i] = a / (float)ex;
  } else if (o < nxz + nyz + nxy) {  // XY plane: mean over z
    const long long i = o - nxz - nyz;
    float a = 0.0f;
    for (int q = 0; q < D.npart[0]; ++q) a += D.part[0][q * nxy + i];
    D.state[0][slot * nxy + i] = a / (float)ez;
  }
}
