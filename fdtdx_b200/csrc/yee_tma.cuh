// TMA-staged Yee E / H half-step kernels for sm_100a.
//
// Same arithmetic, fusion and parameter block as the register-marching kernels in yee_kernels.cuh
// (pad + curl + CPML + material update + PEC/PMC + cold source passes; fdtd/update.py:92-136, 256-354,
// 494-609, 689-930; core/physics/curl.py:227-397), but the field / material tiles of each x plane are
// staged into shared memory by the TMA engine (cp.async.bulk.tensor, 4-D tiled tensor maps over the
// reference's (C,Nx,Ny,Nz) arrays) through an S-deep mbarrier ring.  R warps (one y row x 128 z cells
// each) read their operands with LDS.128 right before use; there is no producer warp: the last warp to
// finish reading a stage (shared-memory arrival counter) re-arms its barrier and issues the tile loads
// of the plane S ahead, so a CTA is exactly 8 warps (2 per scheduler: two CTAs fit the register file).
// What this buys over the register-marching kernels:
//   * bytes in flight are held by the async-copy engine, not by registers x resident warps, so the
//     kernel needs no L2 prefetch instructions and no per-thread 64-bit load addresses;
//   * zero halos (PML / PEC / PMC faces) come from the TMA out-of-bounds fill: the halo tile starts at
//     (k0-4, j0-1) for the E step and ends at (k0+131, j0+R) for the H step, no predicated loads;
//   * the x-neighbour plane of the H step is simply the next ring stage.
// Periodic y / z axes and grids with Nz % 4 != 0 keep using the register-marching kernels.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is resolved at run time in abi.cu)
#include "yee_kernels.cuh"

#define FDTDX_TMA_TZ 128                 // widest tile: z cells per tile row (one warp x float4)
// Tile shapes: TZ = 128 (a warp owns one row of 128 z cells) or TZ = 64 for grids with Nz <= 64 (a warp
// owns two rows of 64 cells: lanes 0-15 / 16-31), so thin grids fill their lanes.  A CTA is always 8
// warps; it covers RT = 8 * (128 / TZ) rows.  Halo tiles are TZ + 4 wide (one 16-byte pad column group).
#define FDTDX_TMA_XC_MAX 64              // longest x chunk (planes per CTA) the per-plane scalar table holds
// tail of the dynamic shared memory, in floats: [full barriers 2*S <= 16][arrivals <= 8][pad][ztab 3*128][xs XC_MAX]
#define FDTDX_TMA_TAIL_ARR 16
#define FDTDX_TMA_TAIL_ZTAB 32
#define FDTDX_TMA_TAIL_XS (FDTDX_TMA_TAIL_ZTAB + 3 * FDTDX_TMA_TZ)  // sized for the widest tile
#define FDTDX_TMA_TAIL_F (FDTDX_TMA_TAIL_XS + FDTDX_TMA_XC_MAX)
#ifndef FDTDX_TMA_MAXREG
#define FDTDX_TMA_MAXREG 128  // R * 32 = 256 threads x 128 registers: two CTAs per SM
#endif

struct alignas(64) TmaSet {
  CUtensorMap fld_halo;   // neighbour field (H in the E step, E in the H step): box (HZ, R+1, 1, 1)
  CUtensorMap fld_plain;  // field being updated: box (TZ, R, 1, 1)
  CUtensorMap mat_plain;  // inv_eps / inv_mu: box (TZ, R, 1, 1)
  CUtensorMap xhalo;      // H step, x_hi_mode 2: (2,ny,nz) Ey,Ez plane of the next rank, box (HZ, R+1, 1, 1)
};

template <int R, int TZ>  // R warps per CTA
struct TmaGeom {
  static constexpr int WR = 128 / TZ;                            // rows per warp
  static constexpr int RT = R * WR;                              // rows per CTA tile
  static constexpr int HZ = TZ + 4;
  static constexpr int HALO_RAW = HZ * (RT + 1) * 4;             // bytes the TMA writes per halo tile
  static constexpr int HALO_B = (HALO_RAW + 127) / 128 * 128;    // slot size (128-byte aligned)
  static constexpr int PLAIN_B = TZ * RT * 4;
  static constexpr int HALO_F = HALO_B / 4, PLAIN_F = PLAIN_B / 4;
};
// Run-time tile geometry.  TZ > 0: the compile-time shapes above (every member folds to a constant).  TZ == 0
// ("flat" tiles, grids whose rows are neither 64 nor 128 cells: Nz = 4 * lz, 5 <= lz <= 31): a tile row is the whole
// z extent, the CTA's 256 threads are laid over (row, z quad) pairs in row-major order - thread t owns quad t % lz
// of row t / lz - so RT = 256 / lz rows keep (almost) every lane busy whatever the row length (Nz = 68: 255 of 256
// lanes instead of 17 of 32).  The z neighbour across a lane that is not the adjacent quad of the same row (first /
// last quad of a row, first / last lane of a warp) is read from the staged tile instead of by shuffle.
struct TmaRt {
  int TZ, LZ, RT, HZ, HALO_RAW, HALO_B, PLAIN_B, HALO_F, PLAIN_F;
};
template <int R, int TZ>
__host__ __device__ __forceinline__ TmaRt tma_geometry(const int flat_lz) {
  TmaRt g;
  if (TZ > 0) {
    using G = TmaGeom<R, (TZ > 0 ? TZ : 128)>;
    g.TZ = TZ; g.LZ = TZ / 4; g.RT = G::RT; g.HZ = G::HZ; g.HALO_RAW = G::HALO_RAW; g.HALO_B = G::HALO_B; g.PLAIN_B = G::PLAIN_B;
  } else {
    g.LZ = flat_lz; g.TZ = 4 * flat_lz; g.RT = (R * 32) / flat_lz; g.HZ = g.TZ + 4;
    g.HALO_RAW = g.HZ * (g.RT + 1) * 4;
    g.HALO_B = (g.HALO_RAW + 127) / 128 * 128;
    g.PLAIN_B = (g.TZ * g.RT * 4 + 127) / 128 * 128;  // tile slots stay 128-byte aligned; the TMA writes TZ * RT * 4 bytes
  }
  g.HALO_F = g.HALO_B / 4;
  g.PLAIN_F = g.PLAIN_B / 4;
  return g;
}
// dynamic shared memory of a flat-tile launch (host side)
template <int R>
inline int tma_smem_bytes_flat(const int flat_lz, const int nmat, const int S) {
  const TmaRt g = tma_geometry<R, 0>(flat_lz);
  return S * (3 * g.HALO_B + (3 + nmat) * g.PLAIN_B) + FDTDX_TMA_TAIL_F * 4;
}
template <int R, int TZ, int NMAT, int S>
constexpr int tma_smem_bytes() {
  // stages, full barriers, arrival counters (padded to 16 B), z-slab coefficient tables, per-plane x scale
  return S * (3 * TmaGeom<R, TZ>::HALO_B + (3 + NMAT) * TmaGeom<R, TZ>::PLAIN_B) + FDTDX_TMA_TAIL_F * 4;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ Vec<4> lds4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  Vec<4> r;
  r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  return r;
}

extern __shared__ __align__(128) float fdtdx_tma_smem[];

// ------------------------------------------------------------------------------------------------
// CPML for the TMA-staged kernels.  Register budget is what bounds residency here, so the slab state
// is kept minimal: y / z psi are software-pipelined one plane ahead in the same registers that were
// just stored (their latency hides behind a whole plane of work), x-slab psi (a few planes per chunk)
// is loaded where it is used behind an L2 prefetch, and psi base pointers / z coefficients are
// re-read from parameter space / L1 instead of being held.  Arithmetic identical to cpml_axis*.
// ------------------------------------------------------------------------------------------------
struct PmlT {
  bool in_y, zh0, zh1;
  int zm;  // PM == 1: bit e set = cell k0+e lies in this lane's z slab (slabs at any thickness / position)
  int yside, zside;
  int ystride, yoff, zstride, zoff;  // psi index of plane i: i * stride + off
  float ay, by, ky;
};

#define FDTDX_TCPML_SETUP(PSI, AT, BT, KT)                                                                             \
  const AxisPmlDev& px = P.pml[0];                                                                                     \
  const AxisPmlDev& py = P.pml[1];                                                                                     \
  const AxisPmlDev& pz = P.pml[2];                                                                                     \
  PmlT L;                                                                                                              \
  L.in_y = false; L.zh0 = false; L.zh1 = false; L.zm = 0; L.yside = 0; L.zside = 0;                                    \
  bool any_z = false;                                                                                                  \
  Vec<V> psy1, psy2, psz1, psz2;                                                                                       \
  if (PM > 0 && lane_ok) {                                                                                             \
    L.in_y = (j < py.lo_len || j >= py.hi_start);                                                                      \
    if (L.in_y) {                                                                                                      \
      L.yside = (j >= py.hi_start) ? 1 : 0;                                                                            \
      L.ystride = (L.yside ? py.hi_len : py.lo_len) * nz;                                                              \
      L.yoff = (L.yside ? j - py.hi_start : j) * nz + k0;                                                              \
      L.ay = py.AT[j]; L.by = py.BT[j]; L.ky = py.KT[j];                                                               \
      if (!P.simulate) { L.ay = 0.0f; L.by = 1.0f; }                                                                   \
      const long long pidx = (long long)ic0 * L.ystride + L.yoff;                                                      \
      psy1 = ldv<V>(py.PSI[L.yside][0] + pidx);                                                                        \
      psy2 = ldv<V>(py.PSI[L.yside][1] + pidx);                                                                        \
    }                                                                                                                  \
    any_z = (k0 < pz.lo_len || k0 + V > pz.hi_start);                                                                  \
    if (PM == 2 && any_z) {                                                                                            \
      L.zside = (k0 + V > pz.hi_start) ? 1 : 0;                                                                        \
      const int zL = L.zside ? pz.hi_len : pz.lo_len;                                                                  \
      L.zh0 = L.zside ? (k0 >= pz.hi_start) : (k0 < pz.lo_len);                                                        \
      L.zh1 = L.zside ? (k0 + 2 >= pz.hi_start) : (k0 + 2 < pz.lo_len);                                                \
      L.zstride = ny * zL;                                                                                             \
      L.zoff = j * zL + (L.zside ? k0 - pz.hi_start : k0);                                                             \
      const long long pidx = (long long)ic0 * L.zstride + L.zoff;                                                      \
      const float* q1 = pz.PSI[L.zside][0] + pidx;                                                                     \
      const float* q2 = pz.PSI[L.zside][1] + pidx;                                                                     \
      if (L.zh0) {                                                                                                     \
        const float2 t1 = *reinterpret_cast<const float2*>(q1), t2 = *reinterpret_cast<const float2*>(q2);             \
        psz1.v[0] = t1.x; psz1.v[1] = t1.y; psz2.v[0] = t2.x; psz2.v[1] = t2.y;                                        \
      }                                                                                                                \
      if (L.zh1) {                                                                                                     \
        const float2 t1 = *reinterpret_cast<const float2*>(q1 + 2), t2 = *reinterpret_cast<const float2*>(q2 + 2);     \
        psz1.v[2] = t1.x; psz1.v[3] = t1.y; psz2.v[2] = t2.x; psz2.v[3] = t2.y;                                        \
      }                                                                                                                \
    }                                                                                                                  \
    if (PM == 1 && any_z) {                                                                                            \
      /* slabs of any thickness / position: per-cell membership mask, 32-bit psi accesses, same pipelining */       \
      L.zside = (k0 + V > pz.hi_start) ? 1 : 0;                                                                        \
      const int zL = L.zside ? pz.hi_len : pz.lo_len;                                                                  \
      _Pragma("unroll") for (int e = 0; e < V; ++e) {                                                                  \
        const int k = k0 + e;                                                                                          \
        if (L.zside ? (k >= pz.hi_start && k < nz) : (k < pz.lo_len)) L.zm |= 1 << e;                                  \
      }                                                                                                                \
      L.zstride = ny * zL;                                                                                             \
      L.zoff = j * zL + (L.zside ? k0 - pz.hi_start : k0);                                                             \
      const long long pidx = (long long)ic0 * L.zstride + L.zoff;                                                      \
      const float* q1 = pz.PSI[L.zside][0] + pidx;                                                                     \
      const float* q2 = pz.PSI[L.zside][1] + pidx;                                                                     \
      _Pragma("unroll") for (int e = 0; e < V; ++e)                                                                    \
        if (L.zm & (1 << e)) { psz1.v[e] = q1[e]; psz2.v[e] = q2[e]; }                                                 \
    }                                                                                                                  \
  }                                                                                                                    \
  const bool psi_st = P.simulate && P.psi_store;

// L2 prefetch of the x-slab psi lines two planes ahead (x slabs are loaded where they are used)
#define FDTDX_TCPML_PREFETCH(PSI)                                                                                      \
  if (PM > 0 && lane_ok) {                                                                                             \
    const int ip = i + 2;                                                                                              \
    if (ip < ic1 && (ip < px.lo_len || ip >= px.hi_start)) {                                                           \
      const int side = (ip >= px.hi_start) ? 1 : 0;                                                                    \
      const long long pidx = (long long)(side ? ip - px.hi_start : ip) * plane + row;                                  \
      prefetch_l2(px.PSI[side][0] + pidx);                                                                             \
      prefetch_l2(px.PSI[side][1] + pidx);                                                                             \
    }                                                                                                                  \
  }

// one cell of the masked z-slab path (PM == 1): update, store, fetch the next plane's psi
#define FDTDX_TCPML_ZCELL(E_)                                                                                          \
      if (L.zm & (1 << E_)) {                                                                                          \
        if (k1) cpml_axis_v<V, E_, E_ + 1, !REV, true>(az, bz, kz, dzFy, dzFx, psz1, psz2, Kx, Ky);                    \
        else cpml_axis_v<V, E_, E_ + 1, !REV, false>(az, bz, kz, dzFy, dzFx, psz1, psz2, Kx, Ky);                      \
        if (!REV && psi_st) { q1[E_] = psz1.v[E_]; q2[E_] = psz2.v[E_]; }                                              \
        if (i + 1 < ic1) { psz1.v[E_] = q1[E_ + L.zstride]; psz2.v[E_] = q2[E_ + L.zstride]; }                         \
      }

#define FDTDX_TCPML_BLOCK(PSI, AT, BT, KT)                                                                             \
  if (PM > 0) {                                                                                                        \
    if (lane_ok && (i < px.lo_len || i >= px.hi_start)) {                                                              \
      const int side = (i >= px.hi_start) ? 1 : 0;                                                                     \
      const long long pidx = (long long)(side ? i - px.hi_start : i) * plane + row;                                    \
      float* qx1 = px.PSI[side][0] + pidx;                                                                             \
      float* qx2 = px.PSI[side][1] + pidx;                                                                             \
      Vec<V> psx1 = ldv<V>(qx1), psx2 = ldv<V>(qx2);                                                                   \
      float a = px.AT[i], b = px.BT[i];                                                                                \
      const float km1 = px.KT[i];                                                                                      \
      if (!P.simulate) { a = 0.0f; b = 1.0f; }                                                                         \
      if (px.kappa_one) cpml_axis<V, !REV, true>(a, b, km1, dxFz, dxFy, psx1, psx2, Ky, Kz);                           \
      else cpml_axis<V, !REV, false>(a, b, km1, dxFz, dxFy, psx1, psx2, Ky, Kz);                                       \
      if (!REV && psi_st) { stv<V>(qx1, psx1); stv<V>(qx2, psx2); }                                                    \
    }                                                                                                                  \
    if (L.in_y) {                                                                                                      \
      if (py.kappa_one) cpml_axis<V, !REV, true>(L.ay, L.by, L.ky, dyFx, dyFz, psy1, psy2, Kz, Kx);                    \
      else cpml_axis<V, !REV, false>(L.ay, L.by, L.ky, dyFx, dyFz, psy1, psy2, Kz, Kx);                                \
      float* q1 = py.PSI[L.yside][0] + ((long long)i * L.ystride + L.yoff);                                            \
      float* q2 = py.PSI[L.yside][1] + ((long long)i * L.ystride + L.yoff);                                            \
      if (!REV && psi_st) { stv<V>(q1, psy1); stv<V>(q2, psy2); }                                                      \
      if (i + 1 < ic1) { psy1 = ldv<V>(q1 + L.ystride); psy2 = ldv<V>(q2 + L.ystride); }                               \
    }                                                                                                                  \
    if (PM == 2) {                                                                                                     \
      if (L.zh0 || L.zh1) {                                                                                            \
        const Vec<V> az = lds4(ztab + zl * V), bz = lds4(ztab + TZE + zl * V);                                  \
        const Vec<V> kz = lds4(ztab + 2 * TZE + zl * V);                                                                  \
        float* q1 = pz.PSI[L.zside][0] + ((long long)i * L.zstride + L.zoff);                                          \
        float* q2 = pz.PSI[L.zside][1] + ((long long)i * L.zstride + L.zoff);                                          \
        if (L.zh0) {                                                                                                   \
          if (pz.kappa_one) cpml_axis_v<V, 0, 2, !REV, true>(az, bz, kz, dzFy, dzFx, psz1, psz2, Kx, Ky);              \
          else cpml_axis_v<V, 0, 2, !REV, false>(az, bz, kz, dzFy, dzFx, psz1, psz2, Kx, Ky);                          \
          if (!REV && psi_st) {                                                                                        \
            *reinterpret_cast<float2*>(q1) = make_float2(psz1.v[0], psz1.v[1]);                                        \
            *reinterpret_cast<float2*>(q2) = make_float2(psz2.v[0], psz2.v[1]);                                        \
          }                                                                                                            \
          if (i + 1 < ic1) {                                                                                           \
            const float2 t1 = *reinterpret_cast<const float2*>(q1 + L.zstride);                                        \
            const float2 t2 = *reinterpret_cast<const float2*>(q2 + L.zstride);                                        \
            psz1.v[0] = t1.x; psz1.v[1] = t1.y; psz2.v[0] = t2.x; psz2.v[1] = t2.y;                                    \
          }                                                                                                            \
        }                                                                                                              \
        if (L.zh1) {                                                                                                   \
          if (pz.kappa_one) cpml_axis_v<V, 2, 4, !REV, true>(az, bz, kz, dzFy, dzFx, psz1, psz2, Kx, Ky);              \
          else cpml_axis_v<V, 2, 4, !REV, false>(az, bz, kz, dzFy, dzFx, psz1, psz2, Kx, Ky);                          \
          if (!REV && psi_st) {                                                                                        \
            *reinterpret_cast<float2*>(q1 + 2) = make_float2(psz1.v[2], psz1.v[3]);                                    \
            *reinterpret_cast<float2*>(q2 + 2) = make_float2(psz2.v[2], psz2.v[3]);                                    \
          }                                                                                                            \
          if (i + 1 < ic1) {                                                                                           \
            const float2 t1 = *reinterpret_cast<const float2*>(q1 + L.zstride + 2);                                    \
            const float2 t2 = *reinterpret_cast<const float2*>(q2 + L.zstride + 2);                                    \
            psz1.v[2] = t1.x; psz1.v[3] = t1.y; psz2.v[2] = t2.x; psz2.v[3] = t2.y;                                    \
          }                                                                                                            \
        }                                                                                                              \
      }                                                                                                                \
    } else if (L.zm) {                                                                                                 \
      const Vec<V> az = lds4(ztab + zl * V), bz = lds4(ztab + TZE + zl * V);                                    \
      const Vec<V> kz = lds4(ztab + 2 * TZE + zl * V);                                                                    \
      float* q1 = pz.PSI[L.zside][0] + ((long long)i * L.zstride + L.zoff);                                            \
      float* q2 = pz.PSI[L.zside][1] + ((long long)i * L.zstride + L.zoff);                                            \
      const bool k1 = pz.kappa_one;                                                                                    \
      FDTDX_TCPML_ZCELL(0) FDTDX_TCPML_ZCELL(1) FDTDX_TCPML_ZCELL(2) FDTDX_TCPML_ZCELL(3)                              \
    }                                                                                                                  \
  }


#if !defined(FDTDX_BUILD_H)
// ------------------------------------------------------------------------------------------------
// E half-step, TMA-staged.  blockDim = (32, R).
// Stage layout: [Hx halo][Hy halo][Hz halo][Ex][Ey][Ez][inv_eps x TIER]; halo tile origin (k0-4, j0-1).
// ------------------------------------------------------------------------------------------------
template <int TIER, int R, int TZ, bool KONLY = false>
__device__ __forceinline__ void tma_issue_E(const TmaRt& G, const TmaSet& M, uint32_t dst, uint32_t full, int kt0, int j0, int i) {
  // bytes the TMA engine delivers: whole boxes, out-of-bounds parts included (flat tiles: TZ * RT * 4, not the padded slot)
  const int plain_raw = G.TZ * G.RT * 4;
  mbar_expect_tx(full, KONLY ? 3 * G.HALO_RAW : 3 * G.HALO_RAW + (3 + TIER) * plain_raw);
#pragma unroll
  for (int c = 0; c < 3; ++c) tma_load_4d(dst + c * G.HALO_B, &M.fld_halo, kt0 - 4, j0 - 1, i, c, full);
  if (KONLY) return;  // curl-only mode: neither the field being updated nor its material is read
#pragma unroll
  for (int c = 0; c < 3; ++c) tma_load_4d(dst + 3 * G.HALO_B + c * G.PLAIN_B, &M.fld_plain, kt0, j0, i, c, full);
#pragma unroll
  for (int c = 0; c < TIER; ++c) tma_load_4d(dst + 3 * G.HALO_B + (3 + c) * G.PLAIN_B, &M.mat_plain, kt0, j0, i, c, full);
}

// KONLY: write the curl (with its CPML correction) instead of updating E - phase 1 of the full-tensor tier
// (tensor_kernels.cuh): only the three halo tiles of H are staged; P.E is the curl scratch.
template <int TIER, bool REV, bool SIG, bool ADE, bool MET, int PM, int R, int S, int TZ, bool KONLY = false>
__global__ void __maxnreg__(FDTDX_TMA_MAXREG)
    yee_E_tma(const __grid_constant__ StepParams P, const __grid_constant__ TmaSet M, const int t) {
  constexpr int V = 4;
  constexpr bool FLAT = (TZ == 0);
  const TmaRt G = tma_geometry<R, TZ>(P.flat_lz);
  const int HZ = G.HZ, LZ = G.LZ, TZE = G.TZ;  // halo row pitch, lanes per row, tile row length
  const int STAGE_F = 3 * G.HALO_F + (3 + TIER) * G.PLAIN_F;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int kt0 = blockIdx.x * TZE, j0 = blockIdx.y * G.RT;
  // z lane inside the row, row inside the tile (flat tiles: the CTA's threads in row-major (row, quad) order)
  const int tid = warp * 32 + lane;
  const int trow = FLAT ? tid / LZ : warp * (32 / LZ) + lane / LZ;
  const int zl = FLAT ? tid - trow * LZ : lane % LZ;
  const int nz = P.nz, ny = P.ny;
  const int ic0 = P.x_begin + (P.z_reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z) * P.xchunk;
  const int ic1 = min(ic0 + P.xchunk, P.x_end);
  const bool peer_cta = (ic0 == 0);  // the chunk that reads the low neighbour's H plane and owns E[0]
  const int j = j0 + trow;
  const int k0 = kt0 + zl * V;
  const bool lane_ok = (trow < G.RT) && (j < ny) && (k0 < nz);
  const int rows_here = min(G.RT, ny - j0);                      // rows of this tile inside the grid
  const int n_act = min(R, (rows_here * LZ + 31) / 32);          // warps that own a row (the others leave before the loop)
  const bool z_first = (zl == 0) || (FLAT && lane == 0);         // k-1 neighbour not in the previous lane: read the staged tile
  const uint32_t sbase = smem_u32(fdtdx_tma_smem);
  const uint32_t bar_full = sbase + S * STAGE_F * 4;
  int* const arrivals = reinterpret_cast<int*>(fdtdx_tma_smem + S * STAGE_F + FDTDX_TMA_TAIL_ARR);
  float* const ztab = fdtdx_tma_smem + S * STAGE_F + FDTDX_TMA_TAIL_ZTAB;  // a, b, 1/kappa-1 of this tile's z cells
  float* const xs = fdtdx_tma_smem + S * STAGE_F + FDTDX_TMA_TAIL_XS;      // metric x scale of this chunk's planes

  // Programmatic dependent launch: let the next kernel of the stream start scheduling its CTAs into
  // this grid's tail as soon as every CTA of this grid is resident; everything up to the
  // griddepcontrol.wait below touches only constant tables and shared memory.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // CTA-constant tables: z-slab CPML coefficients of this tile and the x metric scale of this chunk
  for (int q = warp * 32 + lane; q < 3 * TZE; q += R * 32) {
    const int tb = q / TZE, k = kt0 + (q - tb * TZE);
    float v = 0.0f;
    if (PM > 0 && k < nz) {
      v = (tb == 0) ? P.pml[2].aE[k] : (tb == 1) ? P.pml[2].bE[k] : P.pml[2].kE[k];
      if (!P.simulate && tb < 2) v = (tb == 0) ? 0.0f : 1.0f;
    }
    ztab[q] = v;
  }
  if (MET) {
    for (int q = warp * 32 + lane; q < ic1 - ic0; q += R * 32) xs[q] = P.sB[0][ic0 + q];
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + s * 8, 1);
      arrivals[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // the previous kernel of the stream (the other half-step) must have completed and flushed from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // reverse pass: update_E_reverse undoes the injection first; it must land in global memory before
  // the first tile load reads it (generic -> async proxy)
  // x-slab neighbour over NVLink: its H[-1] plane of the previous half-step must be final before the
  // register queue below reads it, and its H half-step must have finished reading E[0] before this CTA
  // overwrites it
  if (P.peer_wait != nullptr && peer_cta) peer_wait_cta(P);
  if (!KONLY && REV && P.n_src > 0 && P.src_inline && lane_ok && ic0 < P.src_x1 && ic1 > P.src_x0) {
    src_pass_E<V, TIER>(P, t, true, ic0, ic1, j, k0);
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {  // ring fill: planes ic0 .. ic0+S-1
    for (int s = 0; s < S && ic0 + s < ic1; ++s) tma_issue_E<TIER, R, TZ, KONLY>(G, M, sbase + s * STAGE_F * 4, bar_full + s * 8, kt0, j0, ic0 + s);
  }
  if (warp >= n_act) return;

  // ---------------- consumers ----------------
  const long long plane = (long long)ny * nz;
  const long long N = plane * P.nx;
  const long long row = (long long)j * nz + k0;
  FDTDX_TCPML_SETUP(psiE, aE, bE, kE)

  float sBy = 1.0f;
  Vec<V> sBz;
  if (MET) {
    sBy = lane_ok ? P.sB[1][j] : 1.0f;  // lanes of a partly filled warp may sit past the last row
    sBz = lane_ok ? ldv<V>(P.sB[2] + k0) : zerov<V>();
  }
  long long cell0 = (long long)ic0 * plane + row;  // this thread's first cell of the current plane
  float* pE = (P.E_out ? P.E_out : P.E) + cell0;

  // register queue: Hy, Hz of the previous x plane
  Vec<V> hy_im = zerov<V>(), hz_im = zerov<V>();
  if (lane_ok) {
    if (ic0 > 0) {
      hy_im = ldv<V>(P.H + N + (long long)(ic0 - 1) * plane + row);
      hz_im = ldv<V>(P.H + 2 * N + (long long)(ic0 - 1) * plane + row);
    } else if (P.x_lo_mode == 1) {
      hy_im = ldv<V>(P.H + N + (long long)(P.nx - 1) * plane + row);
      hz_im = ldv<V>(P.H + 2 * N + (long long)(P.nx - 1) * plane + row);
    } else if (P.x_lo_mode == 2) {
      hy_im = ldv<V>(P.haloH + row);
      hz_im = ldv<V>(P.haloH + P.haloH_cs + row);
    }
  }
  const int oh = (trow + 1) * HZ + 4 + zl * V;  // own cells inside a halo tile
  const int op = trow * TZE + zl * V;           // own cells inside a plain tile

  int s = 0;
  uint32_t ph = 0;
  for (int i = ic0; i < ic1; ++i) {
    FDTDX_TCPML_PREFETCH(psiE)
    Vec<V> sg3[3];
    if (SIG) load_sigma<V>(P.sigE, P.sigE_cs, cell0, lane_ok, V, sg3);
    if (ADE && lane_ok && i + 2 < ic1) prefetch_ade(P, N, cell0 + 2 * plane);
    mbar_wait(bar_full + s * 8, ph);
    const float* sb = fdtdx_tma_smem + s * STAGE_F;
    const float* sHx = sb + oh;
    const float* sHy = sb + G.HALO_F + oh;
    const float* sHz = sb + 2 * G.HALO_F + oh;
    const Vec<V> hx = lds4(sHx), hy = lds4(sHy), hz = lds4(sHz);
    const Vec<V> hx_jm = lds4(sHx - HZ), hz_jm = lds4(sHz - HZ);
    // z-neighbour (k-1) of the first element: last element of the previous lane / the tile's pad column
    float hx_l = __shfl_up_sync(0xffffffffu, hx.v[V - 1], 1);
    float hy_l = __shfl_up_sync(0xffffffffu, hy.v[V - 1], 1);
    if (z_first) {
      hx_l = sHx[-1];
      hy_l = sHy[-1];
    }
    float sBx = 1.0f;
    if (MET) sBx = xs[i - ic0];
    Vec<V> Kx, Ky, Kz;
    Vec<V> dxFz, dxFy, dyFx, dyFz, dzFy, dzFx;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float hx_km = (e == 0) ? hx_l : hx.v[e == 0 ? 0 : e - 1];
      const float hy_km = (e == 0) ? hy_l : hy.v[e == 0 ? 0 : e - 1];
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
    hy_im = hy;
    hz_im = hz;
    FDTDX_TCPML_BLOCK(psiE, aE, bE, kE)
    const float* sE = sb + 3 * G.HALO_F + op;
    Vec<V> ex, ey, ez, ie0, ie1, ie2;
    if (!KONLY) {
      ex = lds4(sE); ey = lds4(sE + G.PLAIN_F); ez = lds4(sE + 2 * G.PLAIN_F);
      ie0 = lds4(sE + 3 * G.PLAIN_F);
      if (TIER == 3) {
        ie1 = lds4(sE + 4 * G.PLAIN_F);
        ie2 = lds4(sE + 5 * G.PLAIN_F);
      } else {
        ie1 = ie0;
        ie2 = ie0;
      }
    }
    // all operands of this plane are in registers: the last warp to get here refills the stage
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&arrivals[s], 1) == n_act - 1) {
        arrivals[s] = 0;
        if (i + S < ic1) tma_issue_E<TIER, R, TZ, KONLY>(G, M, sbase + s * STAGE_F * 4, bar_full + s * 8, kt0, j0, i + S);
      }
    }
    if (++s == S) { s = 0; ph ^= 1; }

    Vec<V> o3[3];
    if (KONLY) {
      o3[0] = Kx; o3[1] = Ky; o3[2] = Kz;
    } else {
      const Vec<V> Eo3[3] = {ex, ey, ez}, K3[3] = {Kx, Ky, Kz}, ie3[3] = {ie0, ie1, ie2};
      material_update_E<V, REV, SIG, ADE>(P, N, cell0, lane_ok, V, Eo3, K3, ie3, sg3, o3);
    }
    Vec<V>&o0 = o3[0], &o1 = o3[1], &o2 = o3[2];
    // PEC walls (pec.py:70-77)
    if (!KONLY && P.n_walls > 0 && i >= P.wall_x0[0] && i < P.wall_x1[0]) wall_mask<V>(P, 0, i, j, k0, o0, o1, o2);
    if (lane_ok) {
      stv<V>(pE, o0);
      stv<V>(pE + N, o1);
      stv<V>(pE + 2 * N, o2);
    }
    pE += plane;
    cell0 += plane;
  }
  if (!KONLY && !REV && P.n_src > 0 && P.src_inline && lane_ok && ic0 < P.src_x1 && ic1 > P.src_x0) src_pass_E<V, TIER>(P, t, false, ic0, ic1, j, k0);
  if (P.peer_signal != nullptr && peer_cta) peer_signal_warp(P, 0xffffffffu);
}
#endif  // !FDTDX_BUILD_H

#if !defined(FDTDX_BUILD_E)
// ------------------------------------------------------------------------------------------------
// H half-step, TMA-staged.  Stage layout: [Ex halo][Ey halo][Ez halo][Hx][Hy][Hz][inv_mu x MUT];
// halo tile origin (k0, j0): row +1 is the j+1 neighbour, column +4.. the k+1 neighbour; the x+1
// neighbour plane is the next ring stage (one extra Ey,Ez stage is loaded after the last plane).
// ------------------------------------------------------------------------------------------------
template <int MUT, int R, int TZ, bool KONLY = false>
__device__ __forceinline__ void tma_issue_H(const TmaRt& G, const StepParams& P, const TmaSet& M, uint32_t dst, uint32_t full, int kt0, int j0, int i, int ic1) {
  if (i < ic1) {
    mbar_expect_tx(full, KONLY ? 3 * G.HALO_RAW : 3 * G.HALO_RAW + (3 + MUT) * (G.TZ * G.RT * 4));
#pragma unroll
    for (int c = 0; c < 3; ++c) tma_load_4d(dst + c * G.HALO_B, &M.fld_halo, kt0, j0, i, c, full);
    if (!KONLY) {
#pragma unroll
      for (int c = 0; c < 3; ++c) tma_load_4d(dst + 3 * G.HALO_B + c * G.PLAIN_B, &M.fld_plain, kt0, j0, i, c, full);
#pragma unroll
      for (int c = 0; c < MUT; ++c) tma_load_4d(dst + 3 * G.HALO_B + (3 + c) * G.PLAIN_B, &M.mat_plain, kt0, j0, i, c, full);
    }
  } else {
    // Ey, Ez of the plane after the chunk: in-domain plane, wrap plane, neighbour-rank halo, or
    // (coordinate nx, out of bounds) the zero halo
    mbar_expect_tx(full, 2 * G.HALO_RAW);
    if (i < P.nx || P.x_hi_mode == 0) {
      tma_load_4d(dst + G.HALO_B, &M.fld_halo, kt0, j0, i, 1, full);
      tma_load_4d(dst + 2 * G.HALO_B, &M.fld_halo, kt0, j0, i, 2, full);
    } else if (P.x_hi_mode == 1) {
      tma_load_4d(dst + G.HALO_B, &M.fld_halo, kt0, j0, 0, 1, full);
      tma_load_4d(dst + 2 * G.HALO_B, &M.fld_halo, kt0, j0, 0, 2, full);
    } else {
      // the neighbour's stores were observed through an acquire load in the generic proxy (peer_wait_cta
      // + CTA barrier); order them before this async-proxy read
      asm volatile("fence.proxy.async;" ::: "memory");
      tma_load_4d(dst + G.HALO_B, &M.xhalo, kt0, j0, 0, 0, full);
      tma_load_4d(dst + 2 * G.HALO_B, &M.xhalo, kt0, j0, 1, 0, full);
    }
  }
}

template <int MUT, bool REV, bool SIG, bool MET, int PM, int R, int S, int TZ, bool KONLY = false>
__global__ void __maxnreg__(FDTDX_TMA_MAXREG)
    yee_H_tma(const __grid_constant__ StepParams P, const __grid_constant__ TmaSet M, const int t) {
  constexpr int V = 4;
  constexpr bool FLAT = (TZ == 0);
  const TmaRt G = tma_geometry<R, TZ>(P.flat_lz);
  const int HZ = G.HZ, LZ = G.LZ, TZE = G.TZ;  // halo row pitch, lanes per row, tile row length
  const int STAGE_F = 3 * G.HALO_F + (3 + MUT) * G.PLAIN_F;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int kt0 = blockIdx.x * TZE, j0 = blockIdx.y * G.RT;
  const int tid = warp * 32 + lane;
  const int trow = FLAT ? tid / LZ : warp * (32 / LZ) + lane / LZ;  // row inside the tile, z lane inside the row
  const int zl = FLAT ? tid - trow * LZ : lane % LZ;
  const int nz = P.nz, ny = P.ny;
  const int ic0 = P.x_begin + blockIdx.z * P.xchunk;
  const int ic1 = min(ic0 + P.xchunk, P.x_end);
  const bool peer_cta = (ic1 == P.nx);  // the chunk that reads the high neighbour's E plane and owns H[nx-1]
  const int j = j0 + trow;
  const int k0 = kt0 + zl * V;
  const bool lane_ok = (trow < G.RT) && (j < ny) && (k0 < nz);
  const int rows_here = min(G.RT, ny - j0);
  const int n_act = min(R, (rows_here * LZ + 31) / 32);
  const bool z_last = (zl == LZ - 1) || (FLAT && lane == 31);  // k+1 neighbour not in the next lane: read the staged tile
  const uint32_t sbase = smem_u32(fdtdx_tma_smem);
  const uint32_t bar_full = sbase + S * STAGE_F * 4;
  int* const arrivals = reinterpret_cast<int*>(fdtdx_tma_smem + S * STAGE_F + FDTDX_TMA_TAIL_ARR);
  float* const ztab = fdtdx_tma_smem + S * STAGE_F + FDTDX_TMA_TAIL_ZTAB;  // a, b, 1/kappa-1 of this tile's z cells
  float* const xs = fdtdx_tma_smem + S * STAGE_F + FDTDX_TMA_TAIL_XS;      // metric x scale of this chunk's planes

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // CTA-constant tables: z-slab CPML coefficients of this tile and the x metric scale of this chunk
  for (int q = warp * 32 + lane; q < 3 * TZE; q += R * 32) {
    const int tb = q / TZE, k = kt0 + (q - tb * TZE);
    float v = 0.0f;
    if (PM > 0 && k < nz) {
      v = (tb == 0) ? P.pml[2].aH[k] : (tb == 1) ? P.pml[2].bH[k] : P.pml[2].kH[k];
      if (!P.simulate && tb < 2) v = (tb == 0) ? 0.0f : 1.0f;
    }
    ztab[q] = v;
  }
  if (MET) {
    for (int q = warp * 32 + lane; q < ic1 - ic0; q += R * 32) xs[q] = P.sF[0][ic0 + q];
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + s * 8, 1);
      arrivals[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // x-slab neighbour over NVLink: its E[0] plane of this step must be final before the extra ring stage
  // loads it, and its E half-step must have finished reading H[nx-1] before this CTA overwrites it
  if (P.peer_wait != nullptr && peer_cta) peer_wait_cta(P);
  if (!KONLY && REV && P.n_src > 0 && P.src_inline && lane_ok && ic0 < P.src_x1 && ic1 > P.src_x0) {
    src_pass_H<V, MUT>(P, t, true, ic0, ic1, j, k0);
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {  // ring fill: planes ic0 .. ic0+S-1 (plane ic1 is the Ey,Ez-only stage)
    for (int s = 0; s < S && ic0 + s <= ic1; ++s)
      tma_issue_H<MUT, R, TZ, KONLY>(G, P, M, sbase + s * STAGE_F * 4, bar_full + s * 8, kt0, j0, ic0 + s, ic1);
  }
  if (warp >= n_act) return;

  // ---------------- consumers ----------------
  const long long plane = (long long)ny * nz;
  const long long N = plane * P.nx;
  const long long row = (long long)j * nz + k0;
  FDTDX_TCPML_SETUP(psiH, aH, bH, kH)

  float sFy = 1.0f;
  Vec<V> sFz;
  if (MET) {
    sFy = lane_ok ? P.sF[1][j] : 1.0f;
    sFz = lane_ok ? ldv<V>(P.sF[2] + k0) : zerov<V>();
  }
  long long cell0 = (long long)ic0 * plane + row;
  float* pH = (P.H_out ? P.H_out : P.H) + cell0;
  const int oh = trow * HZ + zl * V;
  const int op = trow * TZE + zl * V;

  int s = 0;
  uint32_t ph = 0;
  for (int i = ic0; i < ic1; ++i) {
    FDTDX_TCPML_PREFETCH(psiH)
    Vec<V> sg3[3];
    if (SIG) load_sigma<V>(P.sigH, P.sigH_cs, cell0, lane_ok, V, sg3);
    int sn = s + 1;
    uint32_t phn = ph;
    if (sn == S) { sn = 0; phn ^= 1; }
    mbar_wait(bar_full + s * 8, ph);
    const float* sb = fdtdx_tma_smem + s * STAGE_F;
    const float* sEx = sb + oh;
    const float* sEy = sb + G.HALO_F + oh;
    const float* sEz = sb + 2 * G.HALO_F + oh;
    const Vec<V> ex = lds4(sEx), ey = lds4(sEy), ez = lds4(sEz);
    const Vec<V> ex_jp = lds4(sEx + HZ), ez_jp = lds4(sEz + HZ);
    float ex_r = __shfl_down_sync(0xffffffffu, ex.v[0], 1);
    float ey_r = __shfl_down_sync(0xffffffffu, ey.v[0], 1);
    if (z_last) {
      ex_r = sEx[V];
      ey_r = sEy[V];
    }
    mbar_wait(bar_full + sn * 8, phn);
    const float* sbn = fdtdx_tma_smem + sn * STAGE_F;
    const Vec<V> ey_n = lds4(sbn + G.HALO_F + oh), ez_n = lds4(sbn + 2 * G.HALO_F + oh);
    float sFx = 1.0f;
    if (MET) sFx = xs[i - ic0];
    Vec<V> Kx, Ky, Kz;
    Vec<V> dxFz, dxFy, dyFx, dyFz, dzFy, dzFx;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float ex_kp = (e == V - 1) ? ex_r : ex.v[e == V - 1 ? e : e + 1];
      const float ey_kp = (e == V - 1) ? ey_r : ey.v[e == V - 1 ? e : e + 1];
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
    FDTDX_TCPML_BLOCK(psiH, aH, bH, kH)
    const float* sH = sb + 3 * G.HALO_F + op;
    Vec<V> hx, hy, hz;
    if (!KONLY) { hx = lds4(sH); hy = lds4(sH + G.PLAIN_F); hz = lds4(sH + 2 * G.PLAIN_F); }
    if (!KONLY && !REV && P.hprev_out != nullptr && lane_ok && hprev_wanted(P, i, j)) {  // H_prev for the detector pass
      float* hp = P.hprev_out + cell0;
      stv<V>(hp, hx);
      stv<V>(hp + N, hy);
      stv<V>(hp + 2 * N, hz);
    }
    Vec<V> im0, im1, im2;
    if (!KONLY && MUT >= 1) {
      im0 = lds4(sH + 3 * G.PLAIN_F);
      if (MUT == 3) {
        im1 = lds4(sH + 4 * G.PLAIN_F);
        im2 = lds4(sH + 5 * G.PLAIN_F);
      } else {
        im1 = im0;
        im2 = im0;
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&arrivals[s], 1) == n_act - 1) {
        arrivals[s] = 0;
        if (i + S <= ic1) tma_issue_H<MUT, R, TZ, KONLY>(G, P, M, sbase + s * STAGE_F * 4, bar_full + s * 8, kt0, j0, i + S, ic1);
      }
    }
    s = sn;
    ph = phn;

    Vec<V> im3[3];
    if (MUT >= 1) { im3[0] = im0; im3[1] = im1; im3[2] = im2; }
    else {
#pragma unroll
      for (int e = 0; e < V; ++e) im3[0].v[e] = P.inv_mu_scalar;
      im3[1] = im3[0]; im3[2] = im3[0];
    }
    Vec<V> o3[3];
    if (KONLY) {
      o3[0] = Kx; o3[1] = Ky; o3[2] = Kz;
    } else {
      const Vec<V> Ho3[3] = {hx, hy, hz}, K3[3] = {Kx, Ky, Kz};
      material_update_H<V, REV, SIG>(P, cell0, lane_ok, V, Ho3, K3, im3, sg3, o3);
    }
    Vec<V>&o0 = o3[0], &o1 = o3[1], &o2 = o3[2];
    if (!KONLY && P.n_walls > 0 && i >= P.wall_x0[1] && i < P.wall_x1[1]) wall_mask<V>(P, 1, i, j, k0, o0, o1, o2);
    if (lane_ok) {
      stv<V>(pH, o0);
      stv<V>(pH + N, o1);
      stv<V>(pH + 2 * N, o2);
    }
    pH += plane;
    cell0 += plane;
  }
  // the extra Ey,Ez stage was consumed as the "next plane" of the last iteration; nothing to release
  if (!KONLY && !REV && P.n_src > 0 && P.src_inline && lane_ok && ic0 < P.src_x1 && ic1 > P.src_x0) src_pass_H<V, MUT>(P, t, false, ic0, ic1, j, k0);
  if (P.peer_signal != nullptr && peer_cta) peer_signal_warp(P, 0xffffffffu);
}
#endif  // !FDTDX_BUILD_E
