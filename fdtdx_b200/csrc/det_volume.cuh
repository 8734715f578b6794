// Detector accumulation for LARGE exact-interpolation regions (DET_VOLUME): full-volume energy videos,
// field videos, volume reductions (update_detector_states, fdtd/update.py:1040-1137; interpolate_fields,
// core/physics/curl.py:86-224; Detector.update of objects/detectors/*.py).
//
// The generic kernels of aux_kernels.cuh spend ~900 instructions per cell (one thread per cell, 64-bit
// index arithmetic, 30 scalar loads, a scalar H_prev copy).  Here a warp owns one (x, y) row of the
// region and a lane four consecutive z cells, and a CTA marches along x:
//   H_prev                  is a plan-wide (3,Nx,Ny,Nz) scratch in the layout of H.  In the forward
//                           direction the H half-step kernel itself stores the H it loaded there
//                           (StepParams::hprev_out) - no copy pass; the time-reversed pass (whose H
//                           kernel first un-injects the sources) and the full-tensor tier fill it with
//                           det_gather_rows_kernel,
//   det_march_kernel        reads every operand row of the co-location stencil with one 128-bit load
//                           (k+1 by shuffle), keeps the x-1 rows of the stencil in registers from the
//                           previous plane (14 row loads per plane instead of 21), evaluates the same
//                           expressions in the same order as colocate_interior / colocate_t
//                           (bit-identical samples), and hands each cell to det_emit - or, for averaged
//                           energy slices, reduces the three means itself (z: warp shuffle; y: the
//                           CTA's 8 rows through shared memory; x: along the march) into small
//                           partial buffers, never writing per-cell energies,
//   det_mean_finish_kernel  folds those partials (z tiles / y tiles / x chunks) in a fixed order.
// Sums are deterministic (fixed geometry, fixed order); they differ from the sequential order of
// det_slice_mean_kernel by float32 rounding only (~1e-7 relative).
#pragma once
#include "aux_kernels.cuh"

#define DETV_ROWS 8     // y rows (warps) per CTA
#define DETV_TZ 128     // z cells per warp pass
#define DETV_XC_MAX 8   // most x planes per CTA (shared-memory staging of the y sums)

// One warp per (c, x, y) row of the region plus its low-side halo: H -> hprev_full at the same index
// (rows outside the grid follow the halo rule in the reader: wrap rows are copied at their wrapped
// index, zero rows are never read).
__global__ void __launch_bounds__(256) det_gather_rows_kernel(const GridDev G, const DetDev* __restrict__ dets, const int t, const int inverse) {
  __shared__ DetDev sD;
  det_stage_descriptor(&sD, dets + blockIdx.y);
  const DetDev& D = sD;
  if (((D.flags & DET_INVERSE) != 0) != (inverse != 0) || !D.on[t] || !(D.flags & DET_VOLUME) || !(D.flags & DET_EXACT)) return;
  const int sx = D.hi[0] - D.lo[0] + 1, sy = D.hi[1] - D.lo[1] + 1;
  const long long rows = 3LL * sx * sy;
  const int lane = threadIdx.x & 31;
  const long long N = (long long)G.nx * G.ny * G.nz;
  for (long long r = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(r % sy);
    const int a = (int)((r / sy) % sx);
    const int c = (int)(r / ((long long)sy * sx));
    int x = D.lo[0] - 1 + a, y = D.lo[1] - 1 + b;
    if (x < 0) { if (G.wrap[0]) x += G.nx; else continue; }
    if (y < 0) { if (G.wrap[1]) y += G.ny; else continue; }
    const long long o = c * N + ((long long)x * G.ny + y) * G.nz;
    const float4* src = reinterpret_cast<const float4*>(G.H + o);
    float4* dst = reinterpret_cast<float4*>(D.hprev_full + o);
    for (int q = lane; q < G.nz / 4; q += 32) dst[q] = src[q];
  }
}

struct Row4 {
  float v[4];
  float nx;  // the element after v[3] (k+1 of the lane's last cell)
};
// A row operand is fetched in two phases so that all row loads of a plane are in flight before the
// first warp shuffle (a shuffle right after each load serialises the round trips: ncu showed 12 us
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

// grid: (z tiles of 128, y tiles of 8 rows, x chunks of xcl planes) of the detector box; blockIdx.z
// also enumerates detectors: z = det * nxc + chunk.  The plane loop has no CTA-wide barrier: warps are
// independent; the slice-mean variant meets once after the loop to fold its 8 rows.
// MODE 1: energy detectors on isotropic / diagonal media - the energy density is evaluated inline
// (metrics.py:55-67, same expressions as det_emit) with 128-bit material loads and one division per
// distinct material component; MODE 0: every other kind goes through det_emit.
__device__ __forceinline__ int detv_mode(const GridDev& G, const DetDev& D) {
  return (D.kind == 1 && G.eps_tier != 9 && G.mu_tier != 9) ? 1 : 0;
}
// NU: non-uniform grid (width-weighted edge averages); uniform grids average with 0.5 * (a + b) and no
// per-call test of the width tables
template <bool NU>
__device__ __forceinline__ float bea_t(const GridDev& G, float cur, float prev, int axis, int idx) {
  if (NU) return bea(G, cur, prev, axis, idx);
  return 0.5f * (cur + prev);
}
template <bool EXACT, int MODE, bool NU>
__global__ void __launch_bounds__(256, 2)
    det_march_kernel(const GridDev G, const DetDev* __restrict__ dets, const int t, const int inverse, const int nxc_max, const int xcl) {
  __shared__ DetDev sD;
  __shared__ __align__(16) float sXZ[MODE == 1 ? DETV_XC_MAX * DETV_ROWS * DETV_TZ : 4];  // [plane][row][z]: energies of this chunk
  const int di = blockIdx.z / nxc_max, xc = blockIdx.z - di * nxc_max;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  {
    pdl_trigger();
    for (int q = tid; q < (int)(sizeof(DetDev) / 4); q += 256) reinterpret_cast<int*>(&sD)[q] = reinterpret_cast<const int*>(dets + di)[q];
    __syncthreads();
    pdl_wait();
  }
  const DetDev& D = sD;
  if (((D.flags & DET_INVERSE) != 0) != (inverse != 0) || !D.on[t] || !(D.flags & DET_VOLUME)) return;
  if (((D.flags & DET_EXACT) != 0) != EXACT || detv_mode(G, D) != MODE || (G.w[0] != nullptr || G.w[1] != nullptr || G.w[2] != nullptr) != NU) return;
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const int lane = threadIdx.x, wrow = threadIdx.y;
  // lanes are aligned to global multiples of 4 in z; cells outside [lo_z, hi_z) are masked at emit
  const int zt0 = (D.lo[2] & ~3) + blockIdx.x * DETV_TZ;
  const int z0 = zt0 + 4 * lane;
  const int ry = blockIdx.y * DETV_ROWS + wrow;
  const int rx0 = xc * xcl, rx1 = min(rx0 + xcl, ex);
  if (zt0 >= D.hi[2] || rx0 >= ex || (int)blockIdx.y * DETV_ROWS >= ey) return;  // CTA-uniform
  const bool fused_mean = MODE == 1 && (D.flags & DET_SLICES) && (D.flags & DET_SLICE_MEAN);
  const bool row_ok = ry < ey;  // warp-uniform
  if (!fused_mean && !row_ok) return;
  const int y = D.lo[1] + ry;
  const bool z_in = z0 < G.nz;                                        // this lane holds grid cells
  const bool lane_in = row_ok && z_in && z0 < D.hi[2] && z0 + 4 > D.lo[2];  // ... and cells of the region
  const bool own_tail = (lane == 31) || (z0 + 4 >= G.nz);             // k+1 of the last cell is not in the next lane
  const long long plane = (long long)G.ny * G.nz, N = plane * G.nx;
  // y-neighbour row (halo rule on the min-y face)
  long long dy = -(long long)G.nz;
  bool y_zero = false;
  if (y == 0) { if (G.wrap[1]) dy = (long long)(G.ny - 1) * G.nz; else y_zero = true; }

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
  const float *E0 = G.E, *E1 = G.E + N, *E2 = G.E + 2 * N, *H0 = G.H, *H1 = G.H + N, *H2 = G.H + 2 * N;
  const float *P0 = D.hprev_full, *P1 = D.hprev_full + N, *P2 = D.hprev_full + 2 * N;

  int zmask = 0;  // cells of this lane that belong to the region
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (lane_in && z0 + e >= D.lo[2] && z0 + e < D.hi[2]) zmask |= 1 << e;
  const float hm_scalar = 0.5f * (1.0f / G.inv_mu_scalar);
  const bool staged = (D.flags & DET_REDUCE) != 0;
  const int slot = D.arr_idx[t];
  const long long n_cells = (long long)ex * ey * ez;
  float* const edst = staged ? D.scratch : D.state[0] + (long long)slot * n_cells;
  // rows of plane x-1, carried from the previous plane of the march (time-centred H already formed)
  Row4 exm, hy_m, hz_mc, hz_mm;
  float accx[4] = {0.f, 0.f, 0.f, 0.f};  // sum over this chunk's planes (YZ mean)

  for (int rx = rx0; rx < rx1 && row_ok; ++rx) {
    const int x = D.lo[0] + rx;
    const long long o_c = ((long long)x * G.ny + y) * G.nz + z0;
    float Es[4][3], Hs[4][3];
    if (EXACT) {
      // phase 1: every load of this plane (and, on the chunk's first plane, the x-1 rows)
      const RowLd l_exc = ld(E0, o_c, false, true);
      const RowLd l_eyc = ld(E1, o_c, false, true), l_eym = ld(E1, o_c + dy, y_zero, true);
      const RowLd l_ezc = ld(E2, o_c, false, false);
      const RowLd n_hxc = ld(H0, o_c, false, false), n_hxm = ld(H0, o_c + dy, y_zero, false);
      const RowLd n_hyc = ld(H1, o_c, false, false);
      const RowLd n_zcc = ld(H2, o_c, false, true), n_zcm = ld(H2, o_c + dy, y_zero, true);
      const RowLd p_hxc = ld(P0, o_c, false, false), p_hxm = ld(P0, o_c + dy, y_zero, false);
      const RowLd p_hyc = ld(P1, o_c, false, false);
      const RowLd p_zcc = ld(P2, o_c, false, true), p_zcm = ld(P2, o_c + dy, y_zero, true);
      if (rx == rx0) {
        long long dx = -plane;
        bool x_zero = false;
        // plane x-1: the previous plane of this array, the wrap plane, the zero halo - or, at the low edge of
        // an x-slab, the neighbour rank's last plane delivered in the (3,ny,nz) xlo buffers
        const float *Ex_m = E0 + o_c, *Hy_m = H1 + o_c, *Hz_m = H2 + o_c, *Py_m = P1 + o_c, *Pz_m = P2 + o_c;
        if (x == 0) {
          if (G.xlo_E != nullptr) {
            const long long o_x = (long long)y * G.nz + z0;
            Ex_m = G.xlo_E + o_x; Hy_m = G.xlo_H + plane + o_x; Hz_m = G.xlo_H + 2 * plane + o_x;
            Py_m = G.xlo_Hp + plane + o_x; Pz_m = G.xlo_Hp + 2 * plane + o_x;
            dx = 0;
          } else if (G.wrap[0]) dx = (long long)(G.nx - 1) * plane;
          else x_zero = true;
        }
        const RowLd l_exm = ld(Ex_m, dx, x_zero, true);
        const RowLd n_hym = ld(Hy_m, dx, x_zero, false), p_hym = ld(Py_m, dx, x_zero, false);
        const RowLd n_zmc = ld(Hz_m, dx, x_zero, true), p_zmc = ld(Pz_m, dx, x_zero, true);
        const RowLd n_zmm = ld(Hz_m, dx + dy, x_zero || y_zero, true), p_zmm = ld(Pz_m, dx + dy, x_zero || y_zero, true);
        exm = row_finish(l_exm, true, own_tail);
        hy_m = detv_hbar(row_finish(p_hym, false, false), row_finish(n_hym, false, false));
        hz_mc = detv_hbar(row_finish(p_zmc, true, own_tail), row_finish(n_zmc, true, own_tail));
        hz_mm = detv_hbar(row_finish(p_zmm, true, own_tail), row_finish(n_zmm, true, own_tail));
      }
      // phase 2: k+1 neighbours by shuffle, time-centred H
      const Row4 exc = row_finish(l_exc, true, own_tail);
      const Row4 eyc = row_finish(l_eyc, true, own_tail), eym = row_finish(l_eym, true, own_tail);
      const Row4 ezc = row_finish(l_ezc, false, false);
      const Row4 hx_c = detv_hbar(row_finish(p_hxc, false, false), row_finish(n_hxc, false, false));
      const Row4 hx_m = detv_hbar(row_finish(p_hxm, false, false), row_finish(n_hxm, false, false));
      const Row4 hy_c = detv_hbar(row_finish(p_hyc, false, false), row_finish(n_hyc, false, false));
      const Row4 hz_cc = detv_hbar(row_finish(p_zcc, true, own_tail), row_finish(n_zcc, true, own_tail));
      const Row4 hz_cm = detv_hbar(row_finish(p_zcm, true, own_tail), row_finish(n_zcm, true, own_tail));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // same expressions, same order as colocate_interior / colocate_t (curl.py:120-222)
        float lo = bea_t<NU>(G, exc.v[e], exm.v[e], 0, x);
        float hi = bea_t<NU>(G, row_at(exc, e + 1), row_at(exm, e + 1), 0, x);
        Es[e][0] = (lo + hi) / 2.0f;
        lo = bea_t<NU>(G, eyc.v[e], eym.v[e], 1, y);
        hi = bea_t<NU>(G, row_at(eyc, e + 1), row_at(eym, e + 1), 1, y);
        Es[e][1] = (lo + hi) / 2.0f;
        Es[e][2] = ezc.v[e];
        Hs[e][0] = bea_t<NU>(G, hx_c.v[e], hx_m.v[e], 1, y);
        Hs[e][1] = bea_t<NU>(G, hy_c.v[e], hy_m.v[e], 0, x);
        const float lx = bea_t<NU>(G, hz_cc.v[e], hz_mc.v[e], 0, x);
        const float lxm = bea_t<NU>(G, hz_cm.v[e], hz_mm.v[e], 0, x);
        const float lxy = bea_t<NU>(G, lx, lxm, 1, y);
        const float hx2 = bea_t<NU>(G, row_at(hz_cc, e + 1), row_at(hz_mc, e + 1), 0, x);
        const float hxm = bea_t<NU>(G, row_at(hz_cm, e + 1), row_at(hz_mm, e + 1), 0, x);
        const float hxy = bea_t<NU>(G, hx2, hxm, 1, y);
        Hs[e][2] = (lxy + hxy) / 2.0f;
      }
      exm = exc; hy_m = hy_c; hz_mc = hz_cc; hz_mm = hz_cm;
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
#pragma unroll
            for (int e = 0; e < 4; ++e) hm[c][e] = hm_scalar;
          } else if (c == 0 || G.mu_tier == 3) {
            const float4 q = *reinterpret_cast<const float4*>(G.mu + c * G.mu_cs + o_c);
            hm[c][0] = 0.5f * (1.0f / q.x); hm[c][1] = 0.5f * (1.0f / q.y); hm[c][2] = 0.5f * (1.0f / q.z); hm[c][3] = 0.5f * (1.0f / q.w);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) hm[c][e] = hm[0][e];
          }
        }
        const long long cell0 = ((long long)rx * ey + ry) * ez + (z0 - D.lo[2]);
        float* const dst = edst;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!(zmask & (1 << e))) continue;
          // eE = sum_c 0.5/inv_eps_c * |E_c|^2 accumulated as (x + y) + z, likewise eH (metrics.py:58-66)
          const float a0 = he[0][e] * (fabsf(Es[e][0]) * fabsf(Es[e][0])), a1 = he[1][e] * (fabsf(Es[e][1]) * fabsf(Es[e][1]));
          const float a2 = he[2][e] * (fabsf(Es[e][2]) * fabsf(Es[e][2]));
          const float b0 = hm[0][e] * (fabsf(Hs[e][0]) * fabsf(Hs[e][0])), b1 = hm[1][e] * (fabsf(Hs[e][1]) * fabsf(Hs[e][1]));
          const float b2 = hm[2][e] * (fabsf(Hs[e][2]) * fabsf(Hs[e][2]));
          const float en = ((a0 + a1) + a2) + ((b0 + b1) + b2);
          ev[e] = en;
          if (!fused_mean) dst[cell0 + e] = en;
        }
      }
    } else if (lane_in) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int z = z0 + e;
        if (!(zmask & (1 << e))) continue;
        const int rz = z - D.lo[2];
        const long long cell = ((long long)rx * ey + ry) * ez + rz;
        det_emit<false>(G, D, t, cell, rx, ry, rz, x, y, z, Es[e], Hs[e], nullptr);
      }
    }
    if (MODE == 1 && fused_mean) {  // CTA-uniform
      // XY mean: sum over z of this row's 128-cell pass -> part[0][ztile][rx][ry]
      float rs = (ev[0] + ev[1]) + (ev[2] + ev[3]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      if (lane == 0) D.part[0][((long long)blockIdx.x * ex + rx) * ey + ry] = rs;
      // YZ mean: sum over the planes of this chunk, in march order
#pragma unroll
      for (int e = 0; e < 4; ++e) accx[e] += ev[e];
      // XZ mean: staged for the fold over this CTA's rows
      *reinterpret_cast<float4*>(&sXZ[((rx - rx0) * DETV_ROWS + wrow) * DETV_TZ + 4 * lane]) = make_float4(ev[0], ev[1], ev[2], ev[3]);
    }
  }
  if (MODE == 1 && fused_mean) {
    if (!row_ok) {  // rows past the region contribute zeros to the y fold
      for (int pl = 0; pl < rx1 - rx0; ++pl)
        *reinterpret_cast<float4*>(&sXZ[(pl * DETV_ROWS + wrow) * DETV_TZ + 4 * lane]) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (lane_in) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int z = z0 + e;
        if (z >= D.lo[2] && z < D.hi[2]) D.part[2][((long long)xc * ey + ry) * ez + (z - D.lo[2])] = accx[e];
      }
    }
    __syncthreads();
    for (int q = tid; q < (rx1 - rx0) * DETV_TZ; q += 256) {
      const int pl = q / DETV_TZ, zc = q - pl * DETV_TZ;
      const int z = zt0 + zc;
      if (z < D.lo[2] || z >= D.hi[2]) continue;
      float a = 0.0f;
#pragma unroll
      for (int w = 0; w < DETV_ROWS; ++w) a += sXZ[(pl * DETV_ROWS + w) * DETV_TZ + zc];
      D.part[1][((long long)blockIdx.y * ex + (rx0 + pl)) * ez + (z - D.lo[2])] = a;
    }
  }
}

// Averaged energy slices: fold the partial sums of det_march_kernel in a fixed order - XY over z tiles,
// XZ over y tiles, YZ over x chunks - and write the three planes.
__global__ void det_mean_finish_kernel(const DetDev* __restrict__ dets, const int di, const int t) {
  pdl_trigger();
  const DetDev D = dets[di];
  pdl_wait();
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const long long nxy = (long long)ex * ey, nxz = (long long)ex * ez, nyz = (long long)ey * ez;
  const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int slot = D.arr_idx[t];
  if (o < nxz) {  // XZ plane: mean over y
    float a = 0.0f;
    for (int q = 0; q < D.npart[1]; ++q) a += D.part[1][q * nxz + o];
    D.state[1][slot * nxz + o] = a / (float)ey;
  } else if (o < nxz + nyz) {  // YZ plane: mean over x
    const long long i = o - nxz;
    float a = 0.0f;
    for (int q = 0; q < D.npart[2]; ++q) a += D.part[2][q * nyz + i];
    D.state[2][slot * nyz + i] = a / (float)ex;
  } else if (o < nxz + nyz + nxy) {  // XY plane: mean over z
    const long long i = o - nxz - nyz;
    float a = 0.0f;
    for (int q = 0; q < D.npart[0]; ++q) a += D.part[0][q * nxy + i];
    D.state[0][slot * nxy + i] = a / (float)ez;
  }
}
