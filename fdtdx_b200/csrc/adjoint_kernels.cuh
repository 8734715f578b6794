// Adjoint (VJP) of one forward Yee step, diagonal material tier: the body of reversible_fdtd's
// backward loop (fdtd/fdtd.py:215-251: `backward` then jax.vjp(forward_single_args_wrapper) at the
// reconstructed state).  SURVEY.md Appendix B writes the transposes out; this file implements them:
//
//   det_adjoint_kernel : cotangent of the detector samples through the co-location stencil
//                        (transpose of curl.py:120-222) into lambda_E', lambda_H', lambda_H_prev
//   adj_local_kernel   : per cell, transpose of the material update (update.py:298-354 / 736-750):
//                        PEC/PMC mask, u = lambda'/(1+s), lambda_in = (1-s) u, the material
//                        gradient g += u (+-cK - a F - a F'), lambda_K = +-c inv u, and the CPML
//                        transpose (lambda_psi <- b tot, lambda_d = ... ) -> 6 derivative cotangents
//   adj_gather_kernel  : transpose of the finite differences (with the halo rule transposed) into
//                        the other field's cotangent
//
// The curl K is evaluated on the reconstructed fields with the frozen final psi, exactly as the
// reference's VJP does (Appendix C.1).  One thread per cell; cotangents are carried in float32.
#pragma once
#include "aux_kernels.cuh"
#include "common.cuh"

#ifndef FDTDX_ADJ_MIN_CTAS
#define FDTDX_ADJ_MIN_CTAS 2
#endif

struct AdjParams {
  int nx, ny, nz;
  int wrap[3];
  float cour, eta0, mat_scalar;
  int is_E;            // 1: adjoint of the E half-step (curl of H, backward differences)
  const float* F;      // this half-step's field before the update (E_t or H_t)
  const float* G;      // the field whose curl drives it (H_t for E, E_{t+1} for H)
  const float* mat;    // inv_eps / inv_mu, tier 1|3 (nullptr: scalar)
  long long mat_cs;
  int mat_tier;        // 0 scalar, 1, 3
  const float* sig;    // conductivity (pre-scaled) or nullptr
  long long sig_cs;
  const float* sc[3];  // metric scales of the stencil or nullptr
  AxisPmlDev pml[3];   // primal psi (frozen)
  float* lam_psi[3][2][2];  // cotangent of psi per axis / side / which (same shapes as psi) or nullptr
  float* lam_psi_new[3][2][2];  // fused kernel only: where the updated psi cotangents go (ping-pong partner of lam_psi)
  int xchunk;                   // fused kernel only: x planes per CTA
  int flat_lz;                  // fused kernels only: thin rows (Nz = 4 * flat_lz <= 124) - threads laid over (row, z quad) pairs; 0: warp = row
  int psi_vec;                  // fused kernel only: every psi / psi-cotangent buffer is 16-byte aligned (128-bit rows)
  float* lamF;         // in: cotangent of the updated field; out: cotangent of the input field
  float* lamG;         // cotangent of the other field (accumulated by the gather)
  const float* lam_extra;  // extra input-field cotangent (detector H_prev part) or nullptr
  float* ld;           // scratch (6,N): cotangents of the six derivatives
  float* g_mat;        // gradient accumulator (tier comps, N) or nullptr
  int n_walls;
  const WallDev* walls;
  // ADE (E half-step only; update.py:316-350): primal polarisations of the step's input state, the
  // recurrence coefficients, the cotangents of P / P_prev (updated in place) and the coefficient gradients
  int n_poles, has_c4;
  long long c_cs;            // coefficient component stride (0: isotropic coefficients)
  const float *Pc, *Pq;      // (n_poles, 3, N): P_curr, P_prev
  const float* cf[4];        // c1..c4 (n_poles, 1|3, N); cf[3] nullptr without CCPR poles
  float *lamP, *lamQ;        // (n_poles, 3, N)
  float* g_c[4];             // gradient accumulators shaped like the coefficients, or nullptr
};

__device__ __forceinline__ float a_at(const AdjParams& P, const float* F, int c, int x, int y, int z) {
  if (x < 0) { if (P.wrap[0]) x += P.nx; else return 0.0f; }
  if (x >= P.nx) { if (P.wrap[0]) x -= P.nx; else return 0.0f; }
  if (y < 0) { if (P.wrap[1]) y += P.ny; else return 0.0f; }
  if (y >= P.ny) { if (P.wrap[1]) y -= P.ny; else return 0.0f; }
  if (z < 0) { if (P.wrap[2]) z += P.nz; else return 0.0f; }
  if (z >= P.nz) { if (P.wrap[2]) z -= P.nz; else return 0.0f; }
  const long long N = (long long)P.nx * P.ny * P.nz;
  return F[c * N + ((long long)x * P.ny + y) * P.nz + z];
}

// Per-cell body shared by the scalar and the 4-cells-per-thread kernels: everything after the six
// primal derivatives d[a][c] = d_a G_c (c != a) are known.  Reads lam (cotangent of the updated field),
// F, the material, psi / lambda_psi of the slabs the cell belongs to; writes lam_in[3], ld[6] and the
// material-gradient contributions (gq[c] per component; the caller accumulates them).
template <bool IS_E, bool PML = true>
__device__ __forceinline__ void adj_local_body(const AdjParams& P, const int x, const int y, const int z, const long long cell, const long long N,
                                               float d[3][3], float lam[3], const float Fv[3], const float mv[3], const float sgv[3],
                                               const float extra[3], float lam_in[3], float ldv6[6], float gq[3]) {
  const int pos[3] = {x, y, z};
  float K[3] = {d[1][2] - d[2][1], d[2][0] - d[0][2], d[0][1] - d[1][0]};
  // CPML state of this cell per axis
  bool inp[3];
  float ca[3], cb[3], ck[3];
  long long pidx[3];
  int side[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const AxisPmlDev& A = P.pml[a];
    const int idx = pos[a];
    inp[a] = PML && (idx < A.lo_len || idx >= A.hi_start);
    side[a] = (idx >= A.hi_start) ? 1 : 0;
    ca[a] = cb[a] = ck[a] = 0.0f;
    pidx[a] = 0;
    if (!inp[a]) continue;
    ca[a] = IS_E ? A.aE[idx] : A.aH[idx];
    cb[a] = IS_E ? A.bE[idx] : A.bH[idx];
    ck[a] = A.kappa_one ? 0.0f : (IS_E ? A.kE[idx] : A.kH[idx]);
    if (a == 0) pidx[a] = ((long long)(side[a] ? x - A.hi_start : x) * P.ny + y) * P.nz + z;
    else if (a == 1) pidx[a] = ((long long)x * (side[a] ? A.hi_len : A.lo_len) + (side[a] ? y - A.hi_start : y)) * P.nz + z;
    else pidx[a] = ((long long)x * P.ny + y) * (side[a] ? A.hi_len : A.lo_len) + (side[a] ? z - A.hi_start : z);
    const int i = (a + 1) % 3, j = (a + 2) % 3;
    const float* q1 = IS_E ? (side[a] ? A.psiE[1][0] : A.psiE[0][0]) : (side[a] ? A.psiH[1][0] : A.psiH[0][0]);
    const float* q2 = IS_E ? (side[a] ? A.psiE[1][1] : A.psiE[0][1]) : (side[a] ? A.psiH[1][1] : A.psiH[0][1]);
    const float p1 = cb[a] * q1[pidx[a]] + ca[a] * d[a][j];
    const float p2 = cb[a] * q2[pidx[a]] + ca[a] * d[a][i];
    K[i] = K[i] - (ck[a] * d[a][j] + p1);
    K[j] = K[j] + (ck[a] * d[a][i] + p2);
  }
  // ---- transpose of the material update ----
  for (int w = 0; w < P.n_walls; ++w) {
    const WallDev W = P.walls[w];
    if (W.kind == (IS_E ? 0 : 1) && x >= W.lo[0] && x < W.hi[0] && y >= W.lo[1] && y < W.hi[1] && z >= W.lo[2] && z < W.hi[2]) {
      if (W.axis != 0) lam[0] = 0.0f;
      if (W.axis != 1) lam[1] = 0.0f;
      if (W.axis != 2) lam[2] = 0.0f;
    }
  }
  float lamK[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float m = mv[c];
    const float Fc = Fv[c];
    const float cK = P.cour * K[c];
    float alpha = 0.0f, sv = 0.0f, u, Fpre;
    if (P.sig) {
      const float sg = sgv[c];
      alpha = IS_E ? ((P.cour * sg) * P.eta0) / 2.0f : ((P.cour * sg) / P.eta0) / 2.0f;
      sv = alpha * m;
      u = lam[c] / (1.0f + sv);
      Fpre = IS_E ? ((1.0f - sv) * Fc + cK * m) / (1.0f + sv) : ((1.0f - sv) * Fc - cK * m) / (1.0f + sv);
    } else {  // sv == 0: the divisions by 1 and the (1 - 0) factors are exact identities
      u = lam[c];
      Fpre = IS_E ? Fc + cK * m : Fc - cK * m;
    }
    gq[c] = IS_E ? u * (cK - alpha * Fc - alpha * Fpre) : u * (-cK - alpha * Fc - alpha * Fpre);
    float lin = (1.0f - sv) * u;
    if (IS_E && P.n_poles > 0) {
      // Transpose of the ADE branch (update.py:316-350), with D = sum_p (P_p - Phat_p), den = 1 + s + m sum c4:
      //   Phat_p = c1 P + c2 Q + c3 E;  E1 = (1-s) E + cK m + m D;  E' = E1 / den;  P' = Phat + c4 E';  Q' = P
      const long long pst = 3 * N, cst = P.c_cs ? 3 * N : N;
      float D = 0.0f, c4sum = 0.0f, gEp = lam[c];
      for (int p = 0; p < P.n_poles; ++p) {
        const long long pi = p * pst + c * N + cell, ci = p * cst + c * P.c_cs + cell;
        const float Pp = P.Pc[pi], Qp = P.Pq[pi];
        const float Phat = (P.cf[0][ci] * Pp + P.cf[1][ci] * Qp) + P.cf[2][ci] * Fc;
        D += Pp - Phat;
        if (P.has_c4) { c4sum += P.cf[3][ci]; gEp += P.cf[3][ci] * P.lamP[pi]; }
      }
      const float den = (1.0f + sv) + m * c4sum;
      const float E1 = ((1.0f - sv) * Fc + cK * m) + m * D;
      const float Epre = E1 / den;
      u = gEp / den;                       // cotangent of E1
      const float g_den = -u * Epre;       // cotangent of den
      const float g_s = g_den - Fc * u;    // cotangent of s = alpha m
      gq[c] = u * (cK + D) + g_den * c4sum + alpha * g_s;
      lin = (1.0f - sv) * u;
      const float gD = m * u;
      for (int p = 0; p < P.n_poles; ++p) {
        const long long pi = p * pst + c * N + cell, ci = p * cst + c * P.c_cs + cell;
        const float Pp = P.Pc[pi], Qp = P.Pq[pi];
        const float lP = P.lamP[pi], lQ = P.lamQ[pi];
        const float gPhat = lP - gD;
        lin += P.cf[2][ci] * gPhat;
        P.lamP[pi] = (gD + P.cf[0][ci] * gPhat) + lQ;
        P.lamQ[pi] = P.cf[1][ci] * gPhat;
        // isotropic coefficients are shared by the three components: accumulate atomically
        if (P.g_c[0]) { if (P.c_cs) P.g_c[0][ci] += Pp * gPhat; else atomicAdd(P.g_c[0] + ci, Pp * gPhat); }
        if (P.g_c[1]) { if (P.c_cs) P.g_c[1][ci] += Qp * gPhat; else atomicAdd(P.g_c[1] + ci, Qp * gPhat); }
        if (P.g_c[2]) { if (P.c_cs) P.g_c[2][ci] += Fc * gPhat; else atomicAdd(P.g_c[2] + ci, Fc * gPhat); }
        if (P.has_c4 && P.g_c[3]) {
          const float v = lP * Epre + g_den * m;
          if (P.c_cs) P.g_c[3][ci] += v; else atomicAdd(P.g_c[3] + ci, v);
        }
      }
    }
    if (P.lam_extra) lin += extra[c];
    lam_in[c] = lin;
    lamK[c] = IS_E ? (P.cour * m) * u : -(P.cour * m) * u;
  }
  // ---- transpose of curl + CPML: cotangent of each derivative d_a G_c ----
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int i = (a + 1) % 3, j = (a + 2) % 3;
    // d1 = d_a G_j enters K_i with sign -, d2 = d_a G_i enters K_j with sign +
    float e1 = -lamK[i], e2 = lamK[j];
    float l1 = e1, l2 = e2;
    if (inp[a]) {
      float* L1 = P.lam_psi[a][side[a]][0];
      float* L2 = P.lam_psi[a][side[a]][1];
      const float t1 = (L1 ? L1[pidx[a]] : 0.0f) + e1;
      const float t2 = (L2 ? L2[pidx[a]] : 0.0f) + e2;
      if (L1) L1[pidx[a]] = cb[a] * t1;
      if (L2) L2[pidx[a]] = cb[a] * t2;
      l1 = e1 * (1.0f + ck[a]) + ca[a] * t1;
      l2 = e2 * (1.0f + ck[a]) + ca[a] * t2;
    }
    ldv6[2 * a + 0] = l1;
    ldv6[2 * a + 1] = l2;
  }
}

__global__ void adj_local_kernel(const AdjParams P) {
  const long long N = (long long)P.nx * P.ny * P.nz;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= N) return;
  const int z = (int)(cell % P.nz);
  const int y = (int)((cell / P.nz) % P.ny);
  const int x = (int)(cell / ((long long)P.nz * P.ny));
  const int pos[3] = {x, y, z};
  // ---- primal derivatives at this cell (same arithmetic as the forward kernels) ----
  float d[3][3];
  const int s = P.is_E ? -1 : +1;
  for (int a = 0; a < 3; ++a) {
    const int ex = (a == 0) ? s : 0, ey = (a == 1) ? s : 0, ez = (a == 2) ? s : 0;
    for (int c = 0; c < 3; ++c) {
      if (c == a) continue;
      const float here = a_at(P, P.G, c, x, y, z), there = a_at(P, P.G, c, x + ex, y + ey, z + ez);
      float v = P.is_E ? (here - there) : (there - here);
      if (P.sc[a]) v = v * P.sc[a][pos[a]];
      d[a][c] = v;
    }
  }
  float lam[3] = {P.lamF[cell], P.lamF[N + cell], P.lamF[2 * N + cell]};
  float Fv[3], mv[3], sgv[3] = {0.f, 0.f, 0.f}, extra[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < 3; ++c) {
    Fv[c] = P.F[c * N + cell];
    mv[c] = P.mat_tier == 0 ? P.mat_scalar : P.mat[(long long)(P.mat_tier == 1 ? 0 : c) * P.mat_cs + cell];
    if (P.sig) sgv[c] = P.sig[c * P.sig_cs + cell];
    if (P.lam_extra) extra[c] = P.lam_extra[c * N + cell];
  }
  float lam_in[3], l6[6], gq[3];
  if (P.is_E) adj_local_body<true>(P, x, y, z, cell, N, d, lam, Fv, mv, sgv, extra, lam_in, l6, gq);
  else adj_local_body<false>(P, x, y, z, cell, N, d, lam, Fv, mv, sgv, extra, lam_in, l6, gq);
  for (int c = 0; c < 3; ++c) P.lamF[c * N + cell] = lam_in[c];
  if (P.g_mat) {
    if (P.mat_tier == 3) { for (int c = 0; c < 3; ++c) P.g_mat[c * N + cell] += gq[c]; }
    else P.g_mat[cell] += (gq[0] + gq[1]) + gq[2];
  }
  for (int q = 0; q < 6; ++q) P.ld[(long long)q * N + cell] = l6[q];
}

// 4 cells per thread, 128-bit accesses (Nz % 4 == 0, 16-byte aligned buffers): blockDim (32, 8),
// grid (z tiles of 128, y tiles of 8, x planes).  Same per-cell arithmetic (adj_local_body).
template <bool IS_E>
__global__ void __launch_bounds__(256, FDTDX_ADJ_MIN_CTAS) adj_local4_kernel(const AdjParams P) {
  constexpr int V = 4;
  const int lane = threadIdx.x;
  const int k0 = (blockIdx.x * 32 + lane) * V;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  const bool active = (k0 < P.nz) && (y < P.ny);
  const long long plane = (long long)P.ny * P.nz, N = plane * P.nx;
  const long long cell0 = (long long)x * plane + (long long)y * P.nz + k0;
  const int s = IS_E ? -1 : +1;
  // neighbour planes / rows (zero halo or wrap); *_ok false: zero
  int xn = x + s, yn = y + s;
  bool xok = true, yok = true;
  if (xn < 0) { if (P.wrap[0]) xn = P.nx - 1; else xok = false; }
  if (xn >= P.nx) { if (P.wrap[0]) xn = 0; else xok = false; }
  if (yn < 0) { if (P.wrap[1]) yn = P.ny - 1; else yok = false; }
  if (yn >= P.ny) { if (P.wrap[1]) yn = 0; else yok = false; }
  Vec<V> g[3], gx[3], gy[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { g[c] = zerov<V>(); gx[c] = zerov<V>(); gy[c] = zerov<V>(); }
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g[c] = ldv<V>(P.G + c * N + cell0);
      if (c != 0 && xok) gx[c] = ldv<V>(P.G + c * N + (long long)xn * plane + (long long)y * P.nz + k0);
      if (c != 1 && yok) gy[c] = ldv<V>(P.G + c * N + (long long)x * plane + (long long)yn * P.nz + k0);
    }
  }
  // z neighbour of the edge element: adjacent lane, or memory / halo at the tile edge
  float gz_edge[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float v = IS_E ? __shfl_up_sync(0xffffffffu, g[c].v[V - 1], 1) : __shfl_down_sync(0xffffffffu, g[c].v[0], 1);
    const bool edge = IS_E ? (lane == 0) : (lane == 31 || k0 + V >= P.nz);
    if (edge) {
      int kz = IS_E ? k0 - 1 : k0 + V;
      bool ok = true;
      if (kz < 0) { if (P.wrap[2]) kz = P.nz - 1; else ok = false; }
      if (kz >= P.nz) { if (P.wrap[2]) kz = 0; else ok = false; }
      v = (ok && active) ? P.G[c * N + (long long)x * plane + (long long)y * P.nz + kz] : 0.0f;
    }
    gz_edge[c] = v;
  }
  if (!active) return;
  const float scx = P.sc[0] ? P.sc[0][x] : 1.0f, scy = P.sc[1] ? P.sc[1][y] : 1.0f;
  Vec<V> scz;
  if (P.sc[2]) scz = ldv<V>(P.sc[2] + k0);
  Vec<V> lamv[3], Fq[3], mq[3], sq[3], xq[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    lamv[c] = ldv<V>(P.lamF + c * N + cell0);
    Fq[c] = ldv<V>(P.F + c * N + cell0);
    if (P.mat_tier != 0 && (c == 0 || P.mat_tier == 3)) mq[c] = ldv<V>(P.mat + (long long)c * P.mat_cs + cell0);
    if (P.sig) sq[c] = ldv<V>(P.sig + c * P.sig_cs + cell0);
    if (P.lam_extra) xq[c] = ldv<V>(P.lam_extra + c * N + cell0);
  }
  // CTA-uniform: does this tile (one x plane, 8 rows, 128 z cells) touch any CPML slab?  Interior tiles
  // run the slab-free instantiation of the body.
  const int y0 = blockIdx.y * blockDim.y, y1 = min(y0 + (int)blockDim.y, P.ny) - 1;
  const int z0 = blockIdx.x * 32 * V, z1 = min(z0 + 32 * V, P.nz) - 1;
  const bool cta_pml = (x < P.pml[0].lo_len || x >= P.pml[0].hi_start) || (y0 < P.pml[1].lo_len || y1 >= P.pml[1].hi_start) ||
                       (z0 < P.pml[2].lo_len || z1 >= P.pml[2].hi_start);
  Vec<V> lin[3], l6[6], gq[3];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    float d[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c != 0) { float v = IS_E ? (g[c].v[e] - gx[c].v[e]) : (gx[c].v[e] - g[c].v[e]); if (P.sc[0]) v = v * scx; d[0][c] = v; }
      if (c != 1) { float v = IS_E ? (g[c].v[e] - gy[c].v[e]) : (gy[c].v[e] - g[c].v[e]); if (P.sc[1]) v = v * scy; d[1][c] = v; }
      if (c != 2) {
        const float there = IS_E ? (e == 0 ? gz_edge[c] : g[c].v[e == 0 ? 0 : e - 1]) : (e == V - 1 ? gz_edge[c] : g[c].v[e == V - 1 ? e : e + 1]);
        float v = IS_E ? (g[c].v[e] - there) : (there - g[c].v[e]);
        if (P.sc[2]) v = v * scz.v[e];
        d[2][c] = v;
      }
    }
    d[0][0] = d[1][1] = d[2][2] = 0.0f;
    float lam[3] = {lamv[0].v[e], lamv[1].v[e], lamv[2].v[e]};
    const float Fv[3] = {Fq[0].v[e], Fq[1].v[e], Fq[2].v[e]};
    float mv[3], sgv[3] = {0.f, 0.f, 0.f}, extra[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mv[c] = P.mat_tier == 0 ? P.mat_scalar : (P.mat_tier == 1 ? mq[0].v[e] : mq[c].v[e]);
      if (P.sig) sgv[c] = sq[c].v[e];
      if (P.lam_extra) extra[c] = xq[c].v[e];
    }
    float lam_in[3], ll[6], gg[3];
    if (cta_pml) adj_local_body<IS_E, true>(P, x, y, k0 + e, cell0 + e, N, d, lam, Fv, mv, sgv, extra, lam_in, ll, gg);
    else adj_local_body<IS_E, false>(P, x, y, k0 + e, cell0 + e, N, d, lam, Fv, mv, sgv, extra, lam_in, ll, gg);
#pragma unroll
    for (int c = 0; c < 3; ++c) { lin[c].v[e] = lam_in[c]; gq[c].v[e] = gg[c]; }
#pragma unroll
    for (int q = 0; q < 6; ++q) l6[q].v[e] = ll[q];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) stv<V>(P.lamF + c * N + cell0, lin[c]);
  if (P.g_mat) {
    if (P.mat_tier == 3) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Vec<V> o = ldv<V>(P.g_mat + c * N + cell0);
#pragma unroll
        for (int e = 0; e < V; ++e) o.v[e] += gq[c].v[e];
        stv<V>(P.g_mat + c * N + cell0, o);
      }
    } else {
      Vec<V> o = ldv<V>(P.g_mat + cell0);
#pragma unroll
      for (int e = 0; e < V; ++e) o.v[e] += (gq[0].v[e] + gq[1].v[e]) + gq[2].v[e];
      stv<V>(P.g_mat + cell0, o);
    }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) stv<V>(P.ld + (long long)q * N + cell0, l6[q]);
}

__global__ void adj_gather_kernel(const AdjParams P) {
  const long long N = (long long)P.nx * P.ny * P.nz;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= N) return;
  const int z = (int)(cell % P.nz);
  const int y = (int)((cell / P.nz) % P.ny);
  const int x = (int)(cell / ((long long)P.nz * P.ny));
  const int pos[3] = {x, y, z};
  const int n[3] = {P.nx, P.ny, P.nz};
  const long long stride[3] = {(long long)P.ny * P.nz, (long long)P.nz, 1};
  for (int c = 0; c < 3; ++c) {
    float acc = 0.0f;
    for (int a = 0; a < 3; ++a) {
      if (a == c) continue;
      const int slot = (c == (a + 2) % 3) ? 2 * a : 2 * a + 1;
      const float* L = P.ld + (long long)slot * N;
      const float s_here = P.sc[a] ? P.sc[a][pos[a]] : 1.0f;
      if (P.is_E) {
        // d[q] = sB[q](G[q] - G[q - e_a])  =>  lamG[p] += sB[p] ld[p] - sB[p+1] ld[p + e_a]
        acc += s_here * L[cell];
        int q = pos[a] + 1;
        bool ok = true;
        if (q >= n[a]) { if (P.wrap[a]) q = 0; else ok = false; }
        if (ok) acc -= (P.sc[a] ? P.sc[a][q] : 1.0f) * L[cell + (long long)(q - pos[a]) * stride[a]];
      } else {
        // d[q] = sF[q](G[q + e_a] - G[q])  =>  lamG[p] += sF[p-1] ld[p - e_a] - sF[p] ld[p]
        acc -= s_here * L[cell];
        int q = pos[a] - 1;
        bool ok = true;
        if (q < 0) { if (P.wrap[a]) q = n[a] - 1; else ok = false; }
        if (ok) acc += (P.sc[a] ? P.sc[a][q] : 1.0f) * L[cell + (long long)(q - pos[a]) * stride[a]];
      }
    }
    P.lamG[c * N + cell] += acc;
  }
}

// 4 cells per thread form of adj_gather_kernel (same accumulation order).
template <bool IS_E>
__global__ void __launch_bounds__(256) adj_gather4_kernel(const AdjParams P) {
  constexpr int V = 4;
  const int lane = threadIdx.x;
  const int k0 = (blockIdx.x * 32 + lane) * V;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  const bool active = (k0 < P.nz) && (y < P.ny);
  const long long plane = (long long)P.ny * P.nz, N = plane * P.nx;
  const long long cell0 = (long long)x * plane + (long long)y * P.nz + k0;
  const int s = IS_E ? +1 : -1;  // the transpose looks the other way
  int xn = x + s, yn = y + s;
  bool xok = true, yok = true;
  if (xn < 0) { if (P.wrap[0]) xn = P.nx - 1; else xok = false; }
  if (xn >= P.nx) { if (P.wrap[0]) xn = 0; else xok = false; }
  if (yn < 0) { if (P.wrap[1]) yn = P.ny - 1; else yok = false; }
  if (yn >= P.ny) { if (P.wrap[1]) yn = 0; else yok = false; }
  // z-derivative cotangents (slots 4, 5) also need the k+-1 neighbour: adjacent lane or tile edge
  Vec<V> Lz[2];
  float Lz_edge[2];
  int kz_edge = IS_E ? k0 + V : k0 - 1;
  bool zok = true;
  if (kz_edge < 0) { if (P.wrap[2]) kz_edge = P.nz - 1; else zok = false; }
  if (kz_edge >= P.nz) { if (P.wrap[2]) kz_edge = 0; else zok = false; }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    Lz[q] = active ? ldv<V>(P.ld + (long long)(4 + q) * N + cell0) : zerov<V>();
    float v = IS_E ? __shfl_down_sync(0xffffffffu, Lz[q].v[0], 1) : __shfl_up_sync(0xffffffffu, Lz[q].v[V - 1], 1);
    const bool edge = IS_E ? (lane == 31 || k0 + V >= P.nz) : (lane == 0);
    if (edge) v = (zok && active) ? P.ld[(long long)(4 + q) * N + (long long)x * plane + (long long)y * P.nz + kz_edge] : 0.0f;
    Lz_edge[q] = v;
  }
  if (!active) return;
  const float scx_h = P.sc[0] ? P.sc[0][x] : 1.0f, scy_h = P.sc[1] ? P.sc[1][y] : 1.0f;
  const float scx_n = (P.sc[0] && xok) ? P.sc[0][xn] : 1.0f, scy_n = (P.sc[1] && yok) ? P.sc[1][yn] : 1.0f;
  Vec<V> scz_h, scz_n;
#pragma unroll
  for (int e = 0; e < V; ++e) { scz_h.v[e] = 1.0f; scz_n.v[e] = 1.0f; }
  if (P.sc[2]) {
    scz_h = ldv<V>(P.sc[2] + k0);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      int kn = k0 + e + s;
      if (kn < 0) kn = P.wrap[2] ? P.nz - 1 : 0;
      if (kn >= P.nz) kn = P.wrap[2] ? 0 : P.nz - 1;
      scz_n.v[e] = P.sc[2][kn];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Vec<V> acc = zerov<V>();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (a == c) continue;
      const int slot = (c == (a + 2) % 3) ? 2 * a : 2 * a + 1;
      Vec<V> here, there = zerov<V>();
      bool ok;
      if (a == 2) {
        here = Lz[slot - 4];
        ok = true;  // per element below
#pragma unroll
        for (int e = 0; e < V; ++e)
          there.v[e] = IS_E ? (e == V - 1 ? Lz_edge[slot - 4] : here.v[e == V - 1 ? e : e + 1]) : (e == 0 ? Lz_edge[slot - 4] : here.v[e == 0 ? 0 : e - 1]);
      } else {
        here = ldv<V>(P.ld + (long long)slot * N + cell0);
        ok = (a == 0) ? xok : yok;
        if (ok) there = ldv<V>(P.ld + (long long)slot * N + (a == 0 ? (long long)xn * plane + (long long)y * P.nz : (long long)x * plane + (long long)yn * P.nz) + k0);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float sh = (a == 0) ? scx_h : (a == 1) ? scy_h : scz_h.v[e];
        const float sn = (a == 0) ? scx_n : (a == 1) ? scy_n : scz_n.v[e];
        bool eok = ok;
        if (a == 2) {
          const int kn = k0 + e + s;
          eok = P.wrap[2] || (kn >= 0 && kn < P.nz);
        }
        if (IS_E) {
          acc.v[e] += sh * here.v[e];
          if (eok) acc.v[e] -= sn * there.v[e];
        } else {
          acc.v[e] -= sh * here.v[e];
          if (eok) acc.v[e] += sn * there.v[e];
        }
      }
    }
    Vec<V> o = ldv<V>(P.lamG + c * N + cell0);
#pragma unroll
    for (int e = 0; e < V; ++e) o.v[e] += acc.v[e];
    stv<V>(P.lamG + c * N + cell0, o);
  }
}

// ---------------------------------------------------------------------------------------------
// detector cotangents
// ---------------------------------------------------------------------------------------------
struct DetAdj {
  const float* cot[4];  // cotangent of the detector state (same layout as the state)
  float* lamE;          // (3,N) cotangent of E'
  float* lamH;          // (3,N) cotangent of H'
  float* lamHprev;      // (3,N) cotangent of the step's input H (through H_bar = (H + H')/2)
  float* g_eps;         // inv_eps gradient (energy detectors) or nullptr
  float* g_mu;          // inv_mu gradient (energy detectors, array-valued inv_mu) or nullptr
  int eps_tier, mu_tier;
};
// a missing cotangent (the loss does not depend on that state leaf) reads as zero
__device__ __forceinline__ float cot_at(const float* c, long long i) { return c ? c[i] : 0.0f; }

__device__ __forceinline__ void a_scatter(const GridDev& G, float* buf, int c, int x, int y, int z, float v) {
  if (x < 0) { if (G.wrap[0]) x += G.nx; else return; }
  if (x >= G.nx) { if (G.wrap[0]) x -= G.nx; else return; }
  if (y < 0) { if (G.wrap[1]) y += G.ny; else return; }
  if (y >= G.ny) { if (G.wrap[1]) y -= G.ny; else return; }
  if (z < 0) { if (G.wrap[2]) z += G.nz; else return; }
  if (z >= G.nz) { if (G.wrap[2]) z -= G.nz; else return; }
  const long long N = (long long)G.nx * G.ny * G.nz;
  atomicAdd(buf + c * N + ((long long)x * G.ny + y) * G.nz + z, v);
}

// weights (w_cur, w_prev) of _backward_edge_average along `axis` at index idx
__device__ __forceinline__ void bea_w(const GridDev& G, int axis, int idx, float* wc, float* wp) {
  const float* w = G.w[axis];
  if (w == nullptr) { *wc = 0.5f; *wp = 0.5f; return; }
  const int gi = idx + (axis == 0 ? G.x_offset : 0);
  const float chw = 0.5f * w[gi], phw = 0.5f * w[gi > 0 ? gi - 1 : 0];
  *wc = phw / (chw + phw);
  *wp = chw / (chw + phw);
}

__global__ void det_adjoint_kernel(const GridDev G, const DetDev D, const DetAdj A, const int t) {
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const long long n = (long long)ex * ey * ez;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= n) return;
  const int rz = (int)(cell % ez);
  const int ry = (int)((cell / ez) % ey);
  const int rx = (int)(cell / ((long long)ez * ey));
  const int x = D.lo[0] + rx, y = D.lo[1] + ry, z = D.lo[2] + rz;
  if (D.flags & DET_CLOSED) {  // only the shell of a closed surface is sampled (det_sample_body)
    const bool fx = ((D.aux >> 0) & 1) && (rx == 0 || rx == ex - 1), fy = ((D.aux >> 1) & 1) && (ry == 0 || ry == ey - 1);
    const bool fz = ((D.aux >> 2) & 1) && (rz == 0 || rz == ez - 1);
    if (!fx && !fy && !fz) return;
  }
  float Es[3], Hs[3];
  colocate(G, D, x, y, z, Es, Hs);
  const int slot = D.arr_idx[t];
  float lE[3] = {0.f, 0.f, 0.f}, lH[3] = {0.f, 0.f, 0.f};
  const float sgn = (D.flags & DET_INVERSE) ? -1.0f : 1.0f;
  if (D.kind == 0 || D.kind == 3) {
    int ci = 0;
    for (int c = 0; c < 6; ++c) {
      if (!(D.comp_mask & (1 << c))) continue;
      float lam = 0.0f;
      if (D.kind == 0) {
        lam = (D.flags & DET_REDUCE) ? A.cot[0][(long long)slot * D.ncomp + ci] * D.weights[cell] / D.wsum
                                     : A.cot[0][((long long)slot * D.ncomp + ci) * n + cell];
      } else {
        const float2* ct = reinterpret_cast<const float2*>(A.cot[0]);
        const float wt = D.window[t];
        for (int f = 0; f < D.nf; ++f) {
          const float2 ph = D.ph_table[(long long)t * D.nf + f];
          float2 g;
          float sc = (D.scale * wt);
          if (D.flags & DET_REDUCE) { g = ct[(long long)f * D.ncomp + ci]; sc = sc * D.weights[cell] / D.wsum; }
          else g = ct[((long long)f * D.ncomp + ci) * n + cell];
          lam += sgn * sc * (g.x * ph.x + g.y * ph.y);
        }
      }
      if (c < 3) lE[c] += lam; else lH[c - 3] += lam;
      ++ci;
    }
  } else if (D.kind == 1) {
    float g;
    if (D.flags & DET_REDUCE) g = A.cot[0][slot] * D.weights[cell];
    else if (D.flags & DET_SLICES) {
      if (D.flags & DET_SLICE_MEAN)
        g = cot_at(A.cot[0], ((long long)slot * ex + rx) * ey + ry) / (float)ez + cot_at(A.cot[1], ((long long)slot * ex + rx) * ez + rz) / (float)ey +
            cot_at(A.cot[2], ((long long)slot * ey + ry) * ez + rz) / (float)ex;
      else
        g = (rz == D.slice_idx[2] ? cot_at(A.cot[0], ((long long)slot * ex + rx) * ey + ry) : 0.0f) +
            (ry == D.slice_idx[1] ? cot_at(A.cot[1], ((long long)slot * ex + rx) * ez + rz) : 0.0f) +
            (rx == D.slice_idx[0] ? cot_at(A.cot[2], ((long long)slot * ey + ry) * ez + rz) : 0.0f);
    } else g = A.cot[0][(long long)slot * n + cell];
    const long long N = (long long)G.nx * G.ny * G.nz;
    const long long gidx = ((long long)x * G.ny + y) * G.nz + z;
    for (int c = 0; c < 3; ++c) {
      const float ie = G.eps[c * G.eps_cs + gidx];
      const float im = G.mu ? G.mu[c * G.mu_cs + gidx] : G.inv_mu_scalar;
      lE[c] = g * Es[c] / ie;
      lH[c] = g * Hs[c] / im;
      if (A.g_eps) atomicAdd(A.g_eps + (A.eps_tier == 1 ? 0 : c) * N + gidx, g * (-0.5f * Es[c] * Es[c] / (ie * ie)));
      // d(energy)/d(inv_mu) (metrics.py:55-67 differentiated w.r.t. both materials)
      if (A.g_mu && G.mu) atomicAdd(A.g_mu + (A.mu_tier == 1 ? 0 : c) * N + gidx, g * (-0.5f * Hs[c] * Hs[c] / (im * im)));
    }
  } else {
    float g[3] = {0.f, 0.f, 0.f};
    const float sg = (D.flags & DET_NEGATIVE) ? -1.0f : 1.0f;
    if (D.flags & DET_CLOSED) {
      // state[slot] = sum over shell cells of (+S_a area on the max face, -S_a area on the min face), metrics.py:120-160
      const int r3[3] = {rx, ry, rz}, e3[3] = {ex, ey, ez};
      for (int a = 0; a < 3; ++a) {
        if (!((D.aux >> a) & 1)) continue;
        float coef = 0.0f;
        if (r3[a] == e3[a] - 1) coef += 1.0f;
        if (r3[a] == 0) coef -= 1.0f;
        g[a] = sg * (A.cot[0][slot] * coef * D.weights[a * n + cell]);
      }
    } else if (D.flags & DET_KEEP_ALL) {
      for (int c = 0; c < 3; ++c)
        g[c] = sg * ((D.flags & DET_REDUCE) ? A.cot[0][(long long)slot * 3 + c] * D.weights[c * n + cell] : A.cot[0][((long long)slot * 3 + c) * n + cell]);
    } else {
      g[D.aux] = sg * ((D.flags & DET_REDUCE) ? A.cot[0][slot] * D.weights[cell] : A.cot[0][(long long)slot * n + cell]);
    }
    // L = g . (E x H): dL/dE = H x g, dL/dH = g x E
    lE[0] = Hs[1] * g[2] - Hs[2] * g[1];
    lE[1] = Hs[2] * g[0] - Hs[0] * g[2];
    lE[2] = Hs[0] * g[1] - Hs[1] * g[0];
    lH[0] = g[1] * Es[2] - g[2] * Es[1];
    lH[1] = g[2] * Es[0] - g[0] * Es[2];
    lH[2] = g[0] * Es[1] - g[1] * Es[0];
  }
  // ---- transpose of the co-location stencil (curl.py:120-222) ----
  if (!(D.flags & DET_EXACT)) {
    for (int c = 0; c < 3; ++c) {
      a_scatter(G, A.lamE, c, x, y, z, lE[c]);
      a_scatter(G, A.lamH, c, x, y, z, lH[c]);
    }
    return;
  }
  float wxc, wxp, wyc, wyp;
  bea_w(G, 0, x, &wxc, &wxp);
  bea_w(G, 1, y, &wyc, &wyp);
  // Ex* = 1/2 [ bea_x(Ex[p], Ex[p-ex]) + bea_x(Ex[p+ez], Ex[p-ex+ez]) ]
  a_scatter(G, A.lamE, 0, x, y, z, 0.5f * wxc * lE[0]);
  a_scatter(G, A.lamE, 0, x - 1, y, z, 0.5f * wxp * lE[0]);
  a_scatter(G, A.lamE, 0, x, y, z + 1, 0.5f * wxc * lE[0]);
  a_scatter(G, A.lamE, 0, x - 1, y, z + 1, 0.5f * wxp * lE[0]);
  a_scatter(G, A.lamE, 1, x, y, z, 0.5f * wyc * lE[1]);
  a_scatter(G, A.lamE, 1, x, y - 1, z, 0.5f * wyp * lE[1]);
  a_scatter(G, A.lamE, 1, x, y, z + 1, 0.5f * wyc * lE[1]);
  a_scatter(G, A.lamE, 1, x, y - 1, z + 1, 0.5f * wyp * lE[1]);
  a_scatter(G, A.lamE, 2, x, y, z, lE[2]);
  // H_bar = (H_prev + H_new) / 2: each stencil weight goes half to lambda_H' and half to lambda_H_prev
  float* tgt[2] = {A.lamH, A.lamHprev};
  for (int q = 0; q < 2; ++q) {
    float* B = tgt[q];
    a_scatter(G, B, 0, x, y, z, 0.5f * wyc * lH[0]);
    a_scatter(G, B, 0, x, y - 1, z, 0.5f * wyp * lH[0]);
    a_scatter(G, B, 1, x, y, z, 0.5f * wxc * lH[1]);
    a_scatter(G, B, 1, x - 1, y, z, 0.5f * wxp * lH[1]);
    for (int dz = 0; dz < 2; ++dz) {
      const float h = 0.5f * 0.5f * lH[2];
      a_scatter(G, B, 2, x, y, z + dz, h * wyc * wxc);
      a_scatter(G, B, 2, x - 1, y, z + dz, h * wyc * wxp);
      a_scatter(G, B, 2, x, y - 1, z + dz, h * wyp * wxc);
      a_scatter(G, B, 2, x - 1, y - 1, z + dz, h * wyp * wxp);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused form of adj_local4 + adj_gather4: ONE x-marching pass per half-step, no derivative-cotangent
// scratch.  The adjoint of a half-step is itself a curl-like stencil on W = +-c m lambda' (transposed
// CPML included), so the cotangents of the six derivatives at the three neighbours a cell's gather needs
// are re-evaluated from lambda' and the material instead of being written to HBM and read back:
//   x neighbour : the plane visited just before (the march runs against the gather direction), carried
//                 in registers; the plane in front of the chunk is evaluated once in a prologue
//   y neighbour : re-evaluated from the adjacent row (its lambda' / material rows are L1 / L2 hits)
//   z neighbour : adjacent element / adjacent lane by shuffle; the lane on a 128-cell tile edge
//                 evaluates the one cell across the edge
// psi cotangents are read from the bound buffers and written to a second set (lam_psi_new): a
// neighbour's re-evaluation must see the value from before this half-step.  lambda' itself is left
// untouched, which is exact whenever the half-step has no conductivity and no extra input-field cotangent (the
// caller adds the detector H_prev part afterwards); PEC / PMC walls zero components of lambda', which every
// reader re-applies (adj_wall_mask) while the owner stores the zeros; other configurations use the two-kernel form.  Same per-cell arithmetic and summation order as
// adj_local_body / adj_gather4_kernel, so the results are bit-identical to them.
// HBM traffic per cell and half-step: lambda' 12 + lamG 24 (+ material 4-12 + primal G 12 + gradient 8
// when a material gradient is accumulated) = 36-68 B, vs 132-156 B for the two-kernel form.
// ---------------------------------------------------------------------------------------------
#ifndef FDTDX_ADJ_FUSED_MIN_CTAS
#define FDTDX_ADJ_FUSED_MIN_CTAS 2
#endif
#ifndef FDTDX_ADJ_GRAD_MIN_CTAS
#define FDTDX_ADJ_GRAD_MIN_CTAS 2
#endif
#define FDTDX_ADJ_MAXW 4  // walls of one kind the fused kernel handles (more: the two-kernel form)
struct AdjAx {
  float ca, cb, ck;
  bool in;
  int side;
};
template <bool IS_E>
__device__ __forceinline__ AdjAx adj_ax(const AxisPmlDev& A, const int idx, const bool ok = true) {
  AdjAx r;
  r.ca = r.cb = r.ck = 0.0f;
  r.side = (idx >= A.hi_start) ? 1 : 0;
  r.in = ok && (idx < A.lo_len || idx >= A.hi_start);
  if (r.in) {
    r.ca = IS_E ? A.aE[idx] : A.aH[idx];
    r.cb = IS_E ? A.bE[idx] : A.bH[idx];
    r.ck = A.kappa_one ? 0.0f : (IS_E ? A.kE[idx] : A.kH[idx]);
  }
  return r;
}
template <int AX>
__device__ __forceinline__ long long adj_pidx(const AdjParams& P, const AdjAx& c, const int x, const int y, const int z) {
  const AxisPmlDev& A = P.pml[AX];
  if (AX == 0) return ((long long)(c.side ? x - A.hi_start : x) * P.ny + y) * P.nz + z;
  if (AX == 1) return ((long long)x * (c.side ? A.hi_len : A.lo_len) + (c.side ? y - A.hi_start : y)) * P.nz + z;
  return ((long long)x * P.ny + y) * (c.side ? A.hi_len : A.lo_len) + (c.side ? z - A.hi_start : z);
}
// cotangents (l1, l2) of the two derivatives along one axis at one cell, from e1 = -lamK_i, e2 = +lamK_j
// (adj_local_body, "transpose of curl + CPML"); STORE: also write the cell's new psi cotangents
template <bool STORE>
__device__ __forceinline__ void adj_ld_pair(const AdjAx& c, const float e1, const float e2, const float* __restrict__ L1, const float* __restrict__ L2,
                                            float* __restrict__ N1, float* __restrict__ N2, const long long pi, float& l1, float& l2) {
  l1 = e1;
  l2 = e2;
  if (c.in) {
    const float t1 = (L1 ? L1[pi] : 0.0f) + e1;
    const float t2 = (L2 ? L2[pi] : 0.0f) + e2;
    if (STORE) {
      if (L1) N1[pi] = c.cb * t1;
      if (L2) N2[pi] = c.cb * t2;
    }
    l1 = e1 * (1.0f + c.ck) + c.ca * t1;
    l2 = e2 * (1.0f + c.ck) + c.ca * t2;
  }
}
template <bool IS_E>
__device__ __forceinline__ float adj_lamK(const AdjParams& P, const float m, const float u) {
  return IS_E ? (P.cour * m) * u : -(P.cour * m) * u;
}
template <bool IS_E>
__device__ __forceinline__ void adj_acc(float& acc, const float sh, const float here, const bool ok, const float sn, const float there) {
  if (IS_E) {
    acc += sh * here;
    if (ok) acc -= sn * there;
  } else {
    acc -= sh * here;
    if (ok) acc += sn * there;
  }
}

// components of lambda' zeroed by the PEC (E half-step) / PMC (H half-step) walls at a cell: bit c set -> component c
// is masked (adj_local_body, "transpose of the material update").  Masking is idempotent, so a neighbour that
// re-reads a cell the owner is zeroing in place gets the same value either way.
template <bool IS_E>
__device__ __forceinline__ int adj_wall_mask(const AdjParams& P, const int x, const int y, const int z) {
  int m = 0;
  for (int w = 0; w < P.n_walls; ++w) {
    const WallDev W = P.walls[w];
    if (W.kind == (IS_E ? 0 : 1) && x >= W.lo[0] && x < W.hi[0] && y >= W.lo[1] && y < W.hi[1] && z >= W.lo[2] && z < W.hi[2])
      m |= (W.axis != 0 ? 1 : 0) | (W.axis != 1 ? 2 : 0) | (W.axis != 2 ? 4 : 0);
  }
  return m;
}

__device__ __forceinline__ void adj_prefetch_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// MT: material tier of this half-step (0 scalar, 1, 3); GRAD: accumulate the material gradient (needs the primal
// curl); ROWS: rows (warps) per CTA.  psi / psi-cotangent rows of x and y slabs move as 128-bit vectors (PV).
template <int V>
__device__ __forceinline__ Vec<V> ldpsi(const float* __restrict__ p, const long long i, const bool pv) {
  Vec<V> r;
  if (p == nullptr) return zerov<V>();
  if (pv) return ldv<V>(p + i);
#pragma unroll
  for (int e = 0; e < V; ++e) r.v[e] = p[i + e];
  return r;
}
template <int V>
__device__ __forceinline__ void stpsi(float* __restrict__ p, const long long i, const Vec<V>& r, const bool pv) {
  if (pv) { stv<V>(p + i, r); return; }
#pragma unroll
  for (int e = 0; e < V; ++e) p[i + e] = r.v[e];
}
// four consecutive cells of an x or y slab (uniform coefficients): adj_ld_pair on vectors
template <bool STORE>
__device__ __forceinline__ void adj_ld_pair4(const AdjAx& c, const Vec<4>& e1, const Vec<4>& e2, const float* __restrict__ L1, const float* __restrict__ L2,
                                             float* __restrict__ N1, float* __restrict__ N2, const long long pi, const bool pv, Vec<4>& l1, Vec<4>& l2) {
  l1 = e1;
  l2 = e2;
  if (c.in) {
    const Vec<4> q1 = ldpsi<4>(L1, pi, pv), q2 = ldpsi<4>(L2, pi, pv);
    Vec<4> n1, n2;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float t1 = q1.v[e] + e1.v[e];
      const float t2 = q2.v[e] + e2.v[e];
      n1.v[e] = c.cb * t1;
      n2.v[e] = c.cb * t2;
      l1.v[e] = e1.v[e] * (1.0f + c.ck) + c.ca * t1;
      l2.v[e] = e2.v[e] * (1.0f + c.ck) + c.ca * t2;
    }
    if (STORE) {
      if (L1) stpsi<4>(N1, pi, n1, pv);
      if (L2) stpsi<4>(N2, pi, n2, pv);
    }
  }
}

template <bool IS_E, bool MET>
__device__ __forceinline__ void adj_acc2(float& acc, const float sh, const float here, const bool ok, const float sn, const float there) {
  // same values as adj_acc: x * 1.0f == x and acc -+ 0.0f == acc exactly, so the uniform-grid form and the select are bit-identical
  const float h = MET ? sh * here : here;
  const float t = ok ? (MET ? sn * there : there) : 0.0f;
  if (IS_E) {
    acc += h;
    acc -= t;
  } else {
    acc -= h;
    acc += t;
  }
}

// 16-byte asynchronous global -> shared copy (LDGSTS): the data of the NEXT plane travels while this plane is computed
__device__ __forceinline__ void adj_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void adj_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void adj_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ Vec<4> adj_lds4(const float4* p) {
  const float4 t = *p;
  Vec<4> r;
  r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  return r;
}
// vectors a thread stages per plane: lambda' (3), material (0 | 1 | 3), y-neighbour lambda' (2) and material (0 | 1 | 2), lamG (3)
template <int MT>
__host__ __device__ constexpr int adj_stage_vecs() { return 3 + (MT == 3 ? 3 : MT) + 2 + (MT == 3 ? 2 : MT) + 3; }
extern __shared__ __align__(16) float4 adj_stage_smem[];

// MT: material tier of this half-step (0 scalar, 1, 3); MET: non-uniform grid (metric scales); ROWS: rows (warps) per CTA.
// ASYNC: every thread stages its own operands of plane i+1 in shared memory with cp.async while it computes plane i
// (two stages of adj_stage_vecs<MT>() x 16 B per thread; no barrier - a thread only ever reads what it staged itself).
template <bool IS_E, int MT, bool MET, int ROWS, bool ASYNC>
__global__ void __launch_bounds__(32 * ROWS, FDTDX_ADJ_FUSED_MIN_CTAS) adj_fused4_kernel(const AdjParams P) {
  constexpr int V = 4;
  constexpr int s = IS_E ? +1 : -1;  // the gather looks this way (transpose of the backward / forward difference)
  // walls of this half-step's kind, staged once per CTA
  __shared__ WallDev s_wall[FDTDX_ADJ_MAXW];
  __shared__ int s_nwall;
  // CPML coefficients of the tile's z cells and of the one cell on either side (x-independent)
  __shared__ float s_za[132], s_zb[132], s_zk[132];
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int n = 0;
    for (int w = 0; w < P.n_walls && n < FDTDX_ADJ_MAXW; ++w) {
      const WallDev W = P.walls[w];
      if (W.kind == (IS_E ? 0 : 1)) s_wall[n++] = W;
    }
    s_nwall = n;
  }
  for (int q = threadIdx.y * 32 + threadIdx.x; q < 130; q += 32 * ROWS) {
    int z = (int)blockIdx.x * 128 - 1 + q;
    if (z < 0) z = P.wrap[2] ? P.nz - 1 : 0;
    if (z >= P.nz) z = P.wrap[2] ? 0 : P.nz - 1;
    const AdjAx c = adj_ax<IS_E>(P.pml[2], z);
    s_za[q] = c.ca;
    s_zb[q] = c.cb;
    s_zk[q] = c.ck;
  }
  __syncthreads();
  // thread -> (row, z quad): a warp per row, or (thin rows, flat_lz quads each) the CTA's threads in row-major order
  // over 256 / flat_lz rows, so that rows shorter than 128 cells still fill the lanes (yee_tma.cuh, TmaRt)
  const int lane = threadIdx.x;
  const int lz = P.flat_lz;
  const int tid = threadIdx.y * 32 + lane;
  const int trow = lz ? tid / lz : (int)threadIdx.y;
  const int rt = lz ? (32 * ROWS) / lz : ROWS;
  const int k0 = lz ? (tid - trow * lz) * V : (blockIdx.x * 32 + lane) * V;
  const bool row_ok = trow < rt && (int)blockIdx.y * rt + trow < P.ny;
  if (__ballot_sync(0xffffffffu, row_ok) == 0) return;  // no CTA-wide barrier after this point
  const int y = min((int)blockIdx.y * rt + trow, P.ny - 1);
  const bool active = row_ok && k0 < P.nz;
  const long long plane = (long long)P.ny * P.nz, N = plane * P.nx;
  const long long row0 = (long long)y * P.nz + k0;
  const int c0 = blockIdx.z * P.xchunk, c1 = min(c0 + P.xchunk, P.nx);
  const int xfirst = IS_E ? c1 - 1 : c0;
  const int nplanes = c1 - c0;
  const bool pv = P.psi_vec != 0;
  // ---- per-thread constants: y / z neighbours, scales, slab membership along y and z ----
  int yn = y + s;
  bool yok = true;
  if (yn < 0) { if (P.wrap[1]) yn = P.ny - 1; else yok = false; }
  if (yn >= P.ny) { if (P.wrap[1]) yn = 0; else yok = false; }
  const long long rown = (long long)yn * P.nz + k0;
  const float scy_h = (MET && P.sc[1]) ? P.sc[1][y] : 1.0f;
  const float scy_n = (MET && P.sc[1] && yok) ? P.sc[1][yn] : 1.0f;
  Vec<V> scz_h, scz_n;
  bool zok[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    scz_h.v[e] = 1.0f;
    scz_n.v[e] = 1.0f;
    const int kn = k0 + e + s;
    zok[e] = P.wrap[2] || (kn >= 0 && kn < P.nz);
  }
  if (MET && P.sc[2] && active) {
    scz_h = ldv<V>(P.sc[2] + k0);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      int kn = k0 + e + s;
      if (kn < 0) kn = P.wrap[2] ? P.nz - 1 : 0;
      if (kn >= P.nz) kn = P.wrap[2] ? 0 : P.nz - 1;
      scz_n.v[e] = P.sc[2][kn];
    }
  }
  // the cell across the tile edge in the gather direction
  const bool a_edge = IS_E ? (lane == 31 || k0 + V >= P.nz) : (lane == 0 || k0 == 0);  // the adjacent lane is not the adjacent quad
  int kza = IS_E ? k0 + V : k0 - 1;
  bool kza_ok = active;
  if (kza < 0) { if (P.wrap[2]) kza = P.nz - 1; else kza_ok = false; }
  if (kza >= P.nz) { if (P.wrap[2]) kza = 0; else kza_ok = false; }
  kza_ok = kza_ok && a_edge;
  AdjAx cy = adj_ax<IS_E>(P.pml[1], y), cyn = adj_ax<IS_E>(P.pml[1], yok ? yn : 0, yok);
  cy.in = cy.in && active;
  cyn.in = cyn.in && active;
  const float cy1k = 1.0f + cy.ck, cyn1k = 1.0f + cyn.ck;
  // y-slab addressing: idx = (x * len + yrel) * nz + k0
  const long long ystep = (long long)(cy.side ? P.pml[1].hi_len : P.pml[1].lo_len) * P.nz, ynstep = (long long)(cyn.side ? P.pml[1].hi_len : P.pml[1].lo_len) * P.nz;
  const long long yrow = (long long)(cy.side ? y - P.pml[1].hi_start : y) * P.nz + k0;
  const long long ynrow = (long long)(cyn.side ? yn - P.pml[1].hi_start : yn) * P.nz + k0;
  // z slabs: which of this thread's cells (bits 0-3), and the cell across the tile edge (bit 4), lie in one.  The
  // own cells of a thread are all on one side (the host keeps plans whose two z slabs share a 4-cell group off this kernel).
  int zbits = 0;
#pragma unroll
  for (int e = 0; e < V; ++e)
    if (active && (k0 + e < P.pml[2].lo_len || k0 + e >= P.pml[2].hi_start)) zbits |= 1 << e;
  if (kza_ok && (kza < P.pml[2].lo_len || kza >= P.pml[2].hi_start)) zbits |= 16;
  const int zside = (k0 + V - 1 >= P.pml[2].hi_start && !(k0 < P.pml[2].lo_len)) ? 1 : 0;
  const int zlen = zside ? P.pml[2].hi_len : P.pml[2].lo_len, zrel = zside ? k0 - P.pml[2].hi_start : k0;
  const int zside_e = (kza >= P.pml[2].hi_start) ? 1 : 0;
  const int zlen_e = zside_e ? P.pml[2].hi_len : P.pml[2].lo_len, zrel_e = zside_e ? kza - P.pml[2].hi_start : kza;
  const int zq0 = k0 - (int)blockIdx.x * 128 + 1;  // index of own cell 0 in the staged z tables
  // walls: which of this thread's cells lie in the (y, z) footprint of wall w - bits 0-3 own cells, 4-7 the y neighbour's,
  // 8 the cell across the tile edge - and the components it zeroes (bits 9-11); the x range is tested per plane
  int wbits[FDTDX_ADJ_MAXW];
  bool may_wall = false;
#pragma unroll
  for (int w = 0; w < FDTDX_ADJ_MAXW; ++w) {
    int b = 0;
    if (w < s_nwall && active) {
      const WallDev& W = s_wall[w];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const bool zin = k0 + e >= W.lo[2] && k0 + e < W.hi[2];
        if (zin && y >= W.lo[1] && y < W.hi[1]) b |= 1 << e;
        if (zin && yok && yn >= W.lo[1] && yn < W.hi[1]) b |= 16 << e;
      }
      if (kza_ok && kza >= W.lo[2] && kza < W.hi[2] && y >= W.lo[1] && y < W.hi[1]) b |= 256;
      if (b) b |= ((W.axis != 0 ? 1 : 0) | (W.axis != 1 ? 2 : 0) | (W.axis != 2 ? 4 : 0)) << 9;
    }
    wbits[w] = b;
    may_wall = may_wall || b != 0;
  }
  // component mask (bit c: component c zeroed) of own cell e / the y neighbour's cell e / the edge cell at plane x
  auto wall_mask = [&](const int x, const int sel) -> int {
    int m = 0;
#pragma unroll
    for (int w = 0; w < FDTDX_ADJ_MAXW; ++w)
      if ((wbits[w] >> sel) & 1)
        if (x >= s_wall[w].lo[0] && x < s_wall[w].hi[0]) m |= (wbits[w] >> 9) & 7;
    return m;
  };
  // material of component c at (plane offset + row offset) as a 4-vector
  auto mat4 = [&](const int c, const long long off) -> Vec<V> {
    Vec<V> r;
    if (MT == 0) {
#pragma unroll
      for (int e = 0; e < V; ++e) r.v[e] = P.mat_scalar;
    } else {
      r = ldv<V>(P.mat + (MT == 1 ? 0 : (long long)c * P.mat_cs) + off);
    }
    return r;
  };
  auto mat1 = [&](const int c, const long long off) -> float { return MT == 0 ? P.mat_scalar : P.mat[(MT == 1 ? 0 : (long long)c * P.mat_cs) + off]; };
  // x / y slab rows (uniform coefficients over the thread's four cells): l = e (1 + ck) + ca (L + e), L' = cb (L + e)
  auto slab4 = [&](const bool store, const float ca, const float cb, const float c1k, const float* __restrict__ L1, const float* __restrict__ L2, float* __restrict__ N1,
                   float* __restrict__ N2, const long long pi, Vec<V>& l1, Vec<V>& l2) {
    const Vec<V> q1 = ldpsi<V>(L1, pi, pv), q2 = ldpsi<V>(L2, pi, pv);
    Vec<V> n1, n2;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float t1 = q1.v[e] + l1.v[e];
      const float t2 = q2.v[e] + l2.v[e];
      n1.v[e] = cb * t1;
      n2.v[e] = cb * t2;
      l1.v[e] = l1.v[e] * c1k + ca * t1;
      l2.v[e] = l2.v[e] * c1k + ca * t2;
    }
    if (store) {
      if (L1) stpsi<V>(N1, pi, n1, pv);
      if (L2) stpsi<V>(N2, pi, n2, pv);
    }
  };

  // ---- carried state ----
  Vec<V> l1xp = zerov<V>(), l2xp = zerov<V>();  // x-derivative cotangents of the plane visited before (x + s)
  bool xok = true;
  float scx_n = 1.0f;
  {  // prologue: the plane in front of the chunk
    int xq = xfirst + s;
    if (xq < 0) { if (P.wrap[0]) xq = P.nx - 1; else xok = false; }
    if (xq >= P.nx) { if (P.wrap[0]) xq = 0; else xok = false; }
    if (xok && active) {
      if (MET && P.sc[0]) scx_n = P.sc[0][xq];
      const long long off = (long long)xq * plane + row0;
      Vec<V> ly = ldv<V>(P.lamF + 1 * N + off), lz = ldv<V>(P.lamF + 2 * N + off);
      if (may_wall) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const int wm = wall_mask(xq, e);
          if (wm & 2) ly.v[e] = 0.0f;
          if (wm & 4) lz.v[e] = 0.0f;
        }
      }
      const Vec<V> my = mat4(1, off), mz = mat4(2, off);
#pragma unroll
      for (int e = 0; e < V; ++e) {  // axis 0: i = 1, j = 2
        l1xp.v[e] = -adj_lamK<IS_E>(P, my.v[e], ly.v[e]);
        l2xp.v[e] = adj_lamK<IS_E>(P, mz.v[e], lz.v[e]);
      }
      const AdjAx cq = adj_ax<IS_E>(P.pml[0], xq);
      if (cq.in)
        slab4(false, cq.ca, cq.cb, 1.0f + cq.ck, cq.side ? P.lam_psi[0][1][0] : P.lam_psi[0][0][0], cq.side ? P.lam_psi[0][1][1] : P.lam_psi[0][0][1], nullptr, nullptr,
              ((long long)(cq.side ? xq - P.pml[0].hi_start : xq) * P.ny + y) * P.nz + k0, l1xp, l2xp);
    }
  }

  long long off = (long long)xfirst * plane + row0;
  const long long doff = IS_E ? -plane : plane;
  constexpr int NV = adj_stage_vecs<MT>(), NM = (MT == 3 ? 3 : MT), NMY = (MT == 3 ? 2 : MT);
  auto slot = [&](const int st, const int v) -> float4* { return adj_stage_smem + ((st * NV + v) * (32 * ROWS) + tid); };
  // stage this thread's operands of the plane at offset o (same guards as the direct loads below)
  auto stage_plane = [&](const int st, const long long o) {
    if (active) {
#pragma unroll
      for (int c = 0; c < 3; ++c) adj_cp_async16(slot(st, c), P.lamF + c * N + o);
      if (MT == 1) adj_cp_async16(slot(st, 3), P.mat + o);
      if (MT == 3) {
#pragma unroll
        for (int c = 0; c < 3; ++c) adj_cp_async16(slot(st, 3 + c), P.mat + (long long)c * P.mat_cs + o);
      }
      if (yok) {
        const long long on = o - row0 + rown;
        adj_cp_async16(slot(st, 3 + NM), P.lamF + 2 * N + on);
        adj_cp_async16(slot(st, 4 + NM), P.lamF + on);
        if (MT == 1) adj_cp_async16(slot(st, 5 + NM), P.mat + on);
        if (MT == 3) {
          adj_cp_async16(slot(st, 5 + NM), P.mat + 2 * P.mat_cs + on);
          adj_cp_async16(slot(st, 6 + NM), P.mat + on);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) adj_cp_async16(slot(st, 5 + NM + NMY + c), P.lamG + c * N + o);
    }
    adj_cp_commit();
  };
  if (ASYNC) stage_plane(0, off);
  for (int it = 0; it < nplanes; ++it, off += doff) {
    const int x = IS_E ? xfirst - it : xfirst + it;
    const float scx_h = (MET && P.sc[0]) ? P.sc[0][x] : 1.0f;
    // L2 prefetch of this thread's lines two planes ahead (holds no registers)
    if (active && it + 2 < nplanes) {
      const long long pf = off + 2 * doff;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        adj_prefetch_l2(P.lamF + c * N + pf);
        adj_prefetch_l2(P.lamG + c * N + pf);
        if (MT == 3) adj_prefetch_l2(P.mat + (long long)c * P.mat_cs + pf);
      }
      if (MT == 1) adj_prefetch_l2(P.mat + pf);
    }
    // ---- loads of this plane ----
    Vec<V> lamv[3], mq[3], lny[2], mny[2], og[3];  // lny: lambda'_z, lambda'_x of the y neighbour row (axis 1: i = 2, j = 0)
#pragma unroll
    for (int c = 0; c < 3; ++c) { lamv[c] = zerov<V>(); mq[c] = zerov<V>(); og[c] = zerov<V>(); }
    lny[0] = lny[1] = mny[0] = mny[1] = zerov<V>();
    if (ASYNC) {
      const int st = it & 1;
      if (it + 1 < nplanes) {  // the next plane's operands start travelling now; this plane's group is the older one
        stage_plane(st ^ 1, off + doff);
        adj_cp_wait<1>();
      } else {
        adj_cp_wait<0>();
      }
      if (active) {
#pragma unroll
        for (int c = 0; c < 3; ++c) lamv[c] = adj_lds4(slot(st, c));
        if (MT == 0) {
#pragma unroll
          for (int e = 0; e < V; ++e) mq[0].v[e] = P.mat_scalar;
        } else {
          mq[0] = adj_lds4(slot(st, 3));
        }
        if (MT == 3) { mq[1] = adj_lds4(slot(st, 4)); mq[2] = adj_lds4(slot(st, 5)); }
        else { mq[1] = mq[0]; mq[2] = mq[0]; }
        if (yok) {
          lny[0] = adj_lds4(slot(st, 3 + NM));
          lny[1] = adj_lds4(slot(st, 4 + NM));
          if (MT == 0) mny[0] = mq[0]; else mny[0] = adj_lds4(slot(st, 5 + NM));
          mny[1] = (MT == 3) ? adj_lds4(slot(st, 6 + NM)) : mny[0];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) og[c] = adj_lds4(slot(st, 5 + NM + NMY + c));
      }
    } else if (active) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        lamv[c] = ldv<V>(P.lamF + c * N + off);
        if (MT == 3 || c == 0) mq[c] = mat4(c, off);
      }
      if (MT != 3) { mq[1] = mq[0]; mq[2] = mq[0]; }
      if (yok) {
        const long long offn = off - row0 + rown;
        lny[0] = ldv<V>(P.lamF + 2 * N + offn);
        lny[1] = ldv<V>(P.lamF + 0 * N + offn);
        mny[0] = mat4(2, offn);
        mny[1] = (MT == 3) ? mat4(0, offn) : mny[0];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) og[c] = ldv<V>(P.lamG + c * N + off);
    }
    if (may_wall) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int wm = wall_mask(x, e);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (wm & (1 << c)) {
            lamv[c].v[e] = 0.0f;
            P.lamF[c * N + off + e] = 0.0f;  // lambda_in of a masked component
          }
        const int wn = wall_mask(x, 4 + e);
        if (wn & 4) lny[0].v[e] = 0.0f;
        if (wn & 1) lny[1].v[e] = 0.0f;
      }
    }
    // ---- own cells: lamK and the six derivative cotangents (new psi cotangents stored) ----
    Vec<V> l1a[3], l2a[3];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float k0v = adj_lamK<IS_E>(P, mq[0].v[e], lamv[0].v[e]);
      const float k1v = adj_lamK<IS_E>(P, mq[1].v[e], lamv[1].v[e]);
      const float k2v = adj_lamK<IS_E>(P, mq[2].v[e], lamv[2].v[e]);
      l1a[0].v[e] = -k1v; l2a[0].v[e] = k2v;   // axis 0: e1 = -lamK_y, e2 = +lamK_z
      l1a[1].v[e] = -k2v; l2a[1].v[e] = k0v;   // axis 1: e1 = -lamK_z, e2 = +lamK_x
      l1a[2].v[e] = -k0v; l2a[2].v[e] = k1v;   // axis 2: e1 = -lamK_x, e2 = +lamK_y
    }
    const long long zrow = (long long)x * P.ny + y;
    if (x < P.pml[0].lo_len || x >= P.pml[0].hi_start) {  // plane-uniform
      const AdjAx cx = adj_ax<IS_E>(P.pml[0], x);
      if (active)
        slab4(true, cx.ca, cx.cb, 1.0f + cx.ck, cx.side ? P.lam_psi[0][1][0] : P.lam_psi[0][0][0], cx.side ? P.lam_psi[0][1][1] : P.lam_psi[0][0][1],
              cx.side ? P.lam_psi_new[0][1][0] : P.lam_psi_new[0][0][0], cx.side ? P.lam_psi_new[0][1][1] : P.lam_psi_new[0][0][1],
              ((long long)(cx.side ? x - P.pml[0].hi_start : x) * P.ny + y) * P.nz + k0, l1a[0], l2a[0]);
    }
    if (cy.in)  // warp-uniform
      slab4(true, cy.ca, cy.cb, cy1k, cy.side ? P.lam_psi[1][1][0] : P.lam_psi[1][0][0], cy.side ? P.lam_psi[1][1][1] : P.lam_psi[1][0][1],
            cy.side ? P.lam_psi_new[1][1][0] : P.lam_psi_new[1][0][0], cy.side ? P.lam_psi_new[1][1][1] : P.lam_psi_new[1][0][1], (long long)x * ystep + yrow, l1a[1], l2a[1]);
    if (zbits & 15) {
      const float* L1 = zside ? P.lam_psi[2][1][0] : P.lam_psi[2][0][0];
      const float* L2 = zside ? P.lam_psi[2][1][1] : P.lam_psi[2][0][1];
      float* N1 = zside ? P.lam_psi_new[2][1][0] : P.lam_psi_new[2][0][0];
      float* N2 = zside ? P.lam_psi_new[2][1][1] : P.lam_psi_new[2][0][1];
      const long long pb = zrow * zlen + zrel;
#pragma unroll
      for (int e = 0; e < V; ++e)
        if ((zbits >> e) & 1) {
          const float ca = s_za[zq0 + e], cb = s_zb[zq0 + e], ck = s_zk[zq0 + e];
          const float t1 = (L1 ? L1[pb + e] : 0.0f) + l1a[2].v[e];
          const float t2 = (L2 ? L2[pb + e] : 0.0f) + l2a[2].v[e];
          if (L1) N1[pb + e] = cb * t1;
          if (L2) N2[pb + e] = cb * t2;
          l1a[2].v[e] = l1a[2].v[e] * (1.0f + ck) + ca * t1;
          l2a[2].v[e] = l2a[2].v[e] * (1.0f + ck) + ca * t2;
        }
    }
    // ---- y neighbour row ----
    Vec<V> l1yn, l2yn;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      l1yn.v[e] = -adj_lamK<IS_E>(P, mny[0].v[e], lny[0].v[e]);
      l2yn.v[e] = adj_lamK<IS_E>(P, mny[1].v[e], lny[1].v[e]);
    }
    if (cyn.in)
      slab4(false, cyn.ca, cyn.cb, cyn1k, cyn.side ? P.lam_psi[1][1][0] : P.lam_psi[1][0][0], cyn.side ? P.lam_psi[1][1][1] : P.lam_psi[1][0][1], nullptr, nullptr,
            (long long)x * ynstep + ynrow, l1yn, l2yn);
    // ---- z neighbour across the lane / tile edge ----
    float l1ze = IS_E ? __shfl_down_sync(0xffffffffu, l1a[2].v[0], 1) : __shfl_up_sync(0xffffffffu, l1a[2].v[V - 1], 1);
    float l2ze = IS_E ? __shfl_down_sync(0xffffffffu, l2a[2].v[0], 1) : __shfl_up_sync(0xffffffffu, l2a[2].v[V - 1], 1);
    if (a_edge) {
      l1ze = 0.0f;
      l2ze = 0.0f;
      if (kza_ok) {
        const long long offe = off - k0 + kza;
        float lx = P.lamF[offe], ly = P.lamF[N + offe];
        if (may_wall) {
          const int wm = wall_mask(x, 8);
          if (wm & 1) lx = 0.0f;
          if (wm & 2) ly = 0.0f;
        }
        l1ze = -adj_lamK<IS_E>(P, mat1(0, offe), lx);
        l2ze = adj_lamK<IS_E>(P, mat1(1, offe), ly);
        if (zbits & 16) {
          const float* L1 = zside_e ? P.lam_psi[2][1][0] : P.lam_psi[2][0][0];
          const float* L2 = zside_e ? P.lam_psi[2][1][1] : P.lam_psi[2][0][1];
          const long long pi = zrow * zlen_e + zrel_e;
          const int zq = IS_E ? zq0 + V : zq0 - 1;
          const float ca = s_za[zq], ck = s_zk[zq];
          const float t1 = (L1 ? L1[pi] : 0.0f) + l1ze;
          const float t2 = (L2 ? L2[pi] : 0.0f) + l2ze;
          l1ze = l1ze * (1.0f + ck) + ca * t1;
          l2ze = l2ze * (1.0f + ck) + ca * t2;
        }
      }
    }
    // ---- gather into the other field's cotangent (adj_gather4_kernel's order: for c, axes a != c ascending) ----
    if (active) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float l1zt = IS_E ? (e == V - 1 ? l1ze : l1a[2].v[e == V - 1 ? e : e + 1]) : (e == 0 ? l1ze : l1a[2].v[e == 0 ? 0 : e - 1]);
        const float l2zt = IS_E ? (e == V - 1 ? l2ze : l2a[2].v[e == V - 1 ? e : e + 1]) : (e == 0 ? l2ze : l2a[2].v[e == 0 ? 0 : e - 1]);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        adj_acc2<IS_E, MET>(a0, scy_h, l1a[1].v[e], yok, scy_n, l1yn.v[e]);               // c = 0, a = 1 (slot 2)
        adj_acc2<IS_E, MET>(a0, scz_h.v[e], l2a[2].v[e], zok[e], scz_n.v[e], l2zt);       // c = 0, a = 2 (slot 5)
        adj_acc2<IS_E, MET>(a1, scx_h, l2a[0].v[e], xok, scx_n, l2xp.v[e]);               // c = 1, a = 0 (slot 1)
        adj_acc2<IS_E, MET>(a1, scz_h.v[e], l1a[2].v[e], zok[e], scz_n.v[e], l1zt);       // c = 1, a = 2 (slot 4)
        adj_acc2<IS_E, MET>(a2, scx_h, l1a[0].v[e], xok, scx_n, l1xp.v[e]);               // c = 2, a = 0 (slot 0)
        adj_acc2<IS_E, MET>(a2, scy_h, l2a[1].v[e], yok, scy_n, l2yn.v[e]);               // c = 2, a = 1 (slot 3)
        og[0].v[e] += a0;
        og[1].v[e] += a1;
        og[2].v[e] += a2;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) stv<V>(P.lamG + c * N + off, og[c]);
    }
    // ---- carry ----
    l1xp = l1a[0];
    l2xp = l2a[0];
    xok = true;
    scx_n = scx_h;
  }
}

// Material gradient of one half-step, g += lambda_in . (+-c K): the primal curl K of G with the frozen psi, evaluated
// like adj_local_body does, times the (already wall-masked) cotangent the fused kernel left in lamF.  Runs after
// adj_fused4_kernel; same x-marching layout (G_y, G_z of the next plane are loaded one plane ahead and carried).
template <bool IS_E, int MT, bool MET, int ROWS>
__global__ void __launch_bounds__(32 * ROWS, FDTDX_ADJ_GRAD_MIN_CTAS) adj_grad4_kernel(const AdjParams P) {
  constexpr int V = 4;
  constexpr int ps = IS_E ? -1 : +1;  // the primal difference looks this way
  __shared__ float s_za[132], s_zb[132], s_zk[132];
  for (int q = threadIdx.y * 32 + threadIdx.x; q < 130; q += 32 * ROWS) {
    int z = (int)blockIdx.x * 128 - 1 + q;
    if (z < 0) z = 0;
    if (z >= P.nz) z = P.nz - 1;
    const AdjAx c = adj_ax<IS_E>(P.pml[2], z);
    s_za[q] = c.ca;
    s_zb[q] = c.cb;
    s_zk[q] = c.ck;
  }
  __syncthreads();
  const int lane = threadIdx.x;
  const int lz = P.flat_lz;  // thread -> (row, z quad) as in adj_fused4_kernel
  const int tid = threadIdx.y * 32 + lane;
  const int trow = lz ? tid / lz : (int)threadIdx.y;
  const int rt = lz ? (32 * ROWS) / lz : ROWS;
  const int k0 = lz ? (tid - trow * lz) * V : (blockIdx.x * 32 + lane) * V;
  const bool row_ok = trow < rt && (int)blockIdx.y * rt + trow < P.ny;
  if (__ballot_sync(0xffffffffu, row_ok) == 0) return;
  const int y = min((int)blockIdx.y * rt + trow, P.ny - 1);
  const bool active = row_ok && k0 < P.nz;
  const long long plane = (long long)P.ny * P.nz, N = plane * P.nx;
  const long long row0 = (long long)y * P.nz + k0;
  const int c0 = blockIdx.z * P.xchunk, c1 = min(c0 + P.xchunk, P.nx);
  const int xfirst = IS_E ? c1 - 1 : c0;  // march towards x + ps
  const int nplanes = c1 - c0;
  const bool pv = P.psi_vec != 0;
  int ynp = y + ps;
  bool ypok = true;
  if (ynp < 0) { if (P.wrap[1]) ynp = P.ny - 1; else ypok = false; }
  if (ynp >= P.ny) { if (P.wrap[1]) ynp = 0; else ypok = false; }
  const long long rownp = (long long)ynp * P.nz + k0;
  const float scy_h = (MET && P.sc[1]) ? P.sc[1][y] : 1.0f;
  Vec<V> scz_h;
#pragma unroll
  for (int e = 0; e < V; ++e) scz_h.v[e] = 1.0f;
  if (MET && P.sc[2] && active) scz_h = ldv<V>(P.sc[2] + k0);
  const bool p_edge = IS_E ? (lane == 0 || k0 == 0) : (lane == 31 || k0 + V >= P.nz);
  int kzp = IS_E ? k0 - 1 : k0 + V;
  bool kzp_ok = active && p_edge;
  if (kzp < 0) { if (P.wrap[2]) kzp = P.nz - 1; else kzp_ok = false; }
  if (kzp >= P.nz) { if (P.wrap[2]) kzp = 0; else kzp_ok = false; }
  AdjAx cy = adj_ax<IS_E>(P.pml[1], y);
  cy.in = cy.in && active;
  const float* Qy1 = IS_E ? (cy.side ? P.pml[1].psiE[1][0] : P.pml[1].psiE[0][0]) : (cy.side ? P.pml[1].psiH[1][0] : P.pml[1].psiH[0][0]);
  const float* Qy2 = IS_E ? (cy.side ? P.pml[1].psiE[1][1] : P.pml[1].psiE[0][1]) : (cy.side ? P.pml[1].psiH[1][1] : P.pml[1].psiH[0][1]);
  const long long ystep = (long long)(cy.side ? P.pml[1].hi_len : P.pml[1].lo_len) * P.nz;
  const long long yrow = (long long)(cy.side ? y - P.pml[1].hi_start : y) * P.nz + k0;
  int zbits = 0;
#pragma unroll
  for (int e = 0; e < V; ++e)
    if (active && (k0 + e < P.pml[2].lo_len || k0 + e >= P.pml[2].hi_start)) zbits |= 1 << e;
  const int zside = (k0 + V - 1 >= P.pml[2].hi_start && !(k0 < P.pml[2].lo_len)) ? 1 : 0;
  const int zlen = zside ? P.pml[2].hi_len : P.pml[2].lo_len, zrel = zside ? k0 - P.pml[2].hi_start : k0;
  const int zq0 = k0 - (int)blockIdx.x * 128 + 1;

  Vec<V> gc1 = zerov<V>(), gc2 = zerov<V>();  // primal G_y, G_z of the plane about to be visited
  long long off = (long long)xfirst * plane + row0;
  const long long doff = IS_E ? -plane : plane;
  if (active) {
    gc1 = ldv<V>(P.G + 1 * N + off);
    gc2 = ldv<V>(P.G + 2 * N + off);
  }
  for (int it = 0; it < nplanes; ++it, off += doff) {
    const int x = IS_E ? xfirst - it : xfirst + it;
    const float scx_h = (MET && P.sc[0]) ? P.sc[0][x] : 1.0f;
    if (active && it + 2 < nplanes) {
      const long long pf = off + 2 * doff;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        adj_prefetch_l2(P.lamF + c * N + pf);
        adj_prefetch_l2(P.G + c * N + pf);
        if (MT == 3) adj_prefetch_l2(P.g_mat + c * N + pf);
      }
      if (MT != 3) adj_prefetch_l2(P.g_mat + pf);
    }
    Vec<V> lamv[3], g0 = zerov<V>(), gx1 = zerov<V>(), gx2 = zerov<V>(), gy0 = zerov<V>(), gy2 = zerov<V>(), o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { lamv[c] = zerov<V>(); o[c] = zerov<V>(); }
    if (active) {
#pragma unroll
      for (int c = 0; c < 3; ++c) lamv[c] = ldv<V>(P.lamF + c * N + off);
      g0 = ldv<V>(P.G + off);
      int xp = x + ps;
      bool xpok = true;
      if (xp < 0) { if (P.wrap[0]) xp = P.nx - 1; else xpok = false; }
      if (xp >= P.nx) { if (P.wrap[0]) xp = 0; else xpok = false; }
      if (xpok) {
        const long long offp = (long long)xp * plane + row0;
        gx1 = ldv<V>(P.G + 1 * N + offp);
        gx2 = ldv<V>(P.G + 2 * N + offp);
      }
      if (ypok) {
        const long long offp = off - row0 + rownp;
        gy0 = ldv<V>(P.G + 0 * N + offp);
        gy2 = ldv<V>(P.G + 2 * N + offp);
      }
      if (MT == 3) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = ldv<V>(P.g_mat + c * N + off);
      } else {
        o[0] = ldv<V>(P.g_mat + off);
      }
    }
    float gze[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const Vec<V>& gcv = (c == 0) ? g0 : gc1;
      float v = IS_E ? __shfl_up_sync(0xffffffffu, gcv.v[V - 1], 1) : __shfl_down_sync(0xffffffffu, gcv.v[0], 1);
      if (p_edge) v = kzp_ok ? P.G[c * N + off - k0 + kzp] : 0.0f;
      gze[c] = v;
    }
    if (active) {
      const bool xin = x < P.pml[0].lo_len || x >= P.pml[0].hi_start;
      AdjAx cx;
      cx.ca = cx.cb = cx.ck = 0.0f;
      Vec<V> qx1 = zerov<V>(), qx2 = zerov<V>(), qy1 = zerov<V>(), qy2 = zerov<V>();
      if (xin) {
        cx = adj_ax<IS_E>(P.pml[0], x);
        const long long xb = ((long long)(cx.side ? x - P.pml[0].hi_start : x) * P.ny + y) * P.nz + k0;
        qx1 = ldpsi<V>(IS_E ? (cx.side ? P.pml[0].psiE[1][0] : P.pml[0].psiE[0][0]) : (cx.side ? P.pml[0].psiH[1][0] : P.pml[0].psiH[0][0]), xb, pv);
        qx2 = ldpsi<V>(IS_E ? (cx.side ? P.pml[0].psiE[1][1] : P.pml[0].psiE[0][1]) : (cx.side ? P.pml[0].psiH[1][1] : P.pml[0].psiH[0][1]), xb, pv);
      }
      if (cy.in) {
        const long long yb = (long long)x * ystep + yrow;
        qy1 = ldpsi<V>(Qy1, yb, pv);
        qy2 = ldpsi<V>(Qy2, yb, pv);
      }
      const float* qz1 = IS_E ? (zside ? P.pml[2].psiE[1][0] : P.pml[2].psiE[0][0]) : (zside ? P.pml[2].psiH[1][0] : P.pml[2].psiH[0][0]);
      const float* qz2 = IS_E ? (zside ? P.pml[2].psiE[1][1] : P.pml[2].psiE[0][1]) : (zside ? P.pml[2].psiH[1][1] : P.pml[2].psiH[0][1]);
      const long long zb = ((long long)x * P.ny + y) * zlen + zrel;
      Vec<V> gq[3];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float v1 = IS_E ? (gc1.v[e] - gx1.v[e]) : (gx1.v[e] - gc1.v[e]);
        float v2 = IS_E ? (gc2.v[e] - gx2.v[e]) : (gx2.v[e] - gc2.v[e]);
        if (MET && P.sc[0]) { v1 = v1 * scx_h; v2 = v2 * scx_h; }
        float w0 = IS_E ? (g0.v[e] - gy0.v[e]) : (gy0.v[e] - g0.v[e]);
        float w2 = IS_E ? (gc2.v[e] - gy2.v[e]) : (gy2.v[e] - gc2.v[e]);
        if (MET && P.sc[1]) { w0 = w0 * scy_h; w2 = w2 * scy_h; }
        const float t0 = IS_E ? (e == 0 ? gze[0] : g0.v[e == 0 ? 0 : e - 1]) : (e == V - 1 ? gze[0] : g0.v[e == V - 1 ? e : e + 1]);
        const float t1 = IS_E ? (e == 0 ? gze[1] : gc1.v[e == 0 ? 0 : e - 1]) : (e == V - 1 ? gze[1] : gc1.v[e == V - 1 ? e : e + 1]);
        float u0 = IS_E ? (g0.v[e] - t0) : (t0 - g0.v[e]);
        float u1 = IS_E ? (gc1.v[e] - t1) : (t1 - gc1.v[e]);
        if (MET && P.sc[2]) { u0 = u0 * scz_h.v[e]; u1 = u1 * scz_h.v[e]; }
        // d[0][1] = v1, d[0][2] = v2, d[1][0] = w0, d[1][2] = w2, d[2][0] = u0, d[2][1] = u1
        float K[3] = {w2 - u1, u0 - v2, v1 - w0};
        if (xin) {
          const float p1 = cx.cb * qx1.v[e] + cx.ca * v2;
          const float p2 = cx.cb * qx2.v[e] + cx.ca * v1;
          K[1] = K[1] - (cx.ck * v2 + p1);
          K[2] = K[2] + (cx.ck * v1 + p2);
        }
        if (cy.in) {
          const float p1 = cy.cb * qy1.v[e] + cy.ca * w0;
          const float p2 = cy.cb * qy2.v[e] + cy.ca * w2;
          K[2] = K[2] - (cy.ck * w0 + p1);
          K[0] = K[0] + (cy.ck * w2 + p2);
        }
        if ((zbits >> e) & 1) {
          const float ca = s_za[zq0 + e], cb = s_zb[zq0 + e], ck = s_zk[zq0 + e];
          const float p1 = cb * qz1[zb + e] + ca * u1;
          const float p2 = cb * qz2[zb + e] + ca * u0;
          K[0] = K[0] - (ck * u1 + p1);
          K[1] = K[1] + (ck * u0 + p2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float cK = P.cour * K[c];
          gq[c].v[e] = IS_E ? lamv[c].v[e] * cK : lamv[c].v[e] * (-cK);
        }
      }
      if (MT == 3) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int e = 0; e < V; ++e) o[c].v[e] += gq[c].v[e];
          stv<V>(P.g_mat + c * N + off, o[c]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) o[0].v[e] += (gq[0].v[e] + gq[1].v[e]) + gq[2].v[e];
        stv<V>(P.g_mat + off, o[0]);
      }
    }
    gc1 = gx1;  // the look-ahead plane becomes the current one
    gc2 = gx2;
  }
}


// lam[box] += extra[box]; extra[box] = 0 (the detector H_prev cotangent lives in an otherwise-zero (3,N) scratch)
__global__ void adj_box_add_clear_kernel(float* __restrict__ lam, float* __restrict__ extra, const int nx, const int ny, const int nz,
                                         const int x0, const int y0, const int z0, const int ex, const int ey, const int ez) {
  const long long n = (long long)ex * ey * ez, N = (long long)nx * ny * nz;
  const bool small = 3 * n < 0x7fffffffLL;  // 32-bit index arithmetic whenever the box allows it
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 3 * n; i += (long long)gridDim.x * blockDim.x) {
    int c, x, yy, z;
    if (small) {
      const unsigned u = (unsigned)i, un = (unsigned)n, uyz = (unsigned)(ey * ez);
      c = (int)(u / un);
      const unsigned r = u - (unsigned)c * un, rx = r / uyz, r2 = r - rx * uyz;
      x = x0 + (int)rx; yy = y0 + (int)(r2 / (unsigned)ez); z = z0 + (int)(r2 % (unsigned)ez);
    } else {
      c = (int)(i / n);
      const long long r = i - c * n;
      z = z0 + (int)(r % ez); yy = y0 + (int)((r / ez) % ey); x = x0 + (int)(r / ((long long)ez * ey));
    }
    const long long g = c * N + ((long long)x * ny + yy) * nz + z;
    const float v = extra[g];
    if (v != 0.0f) {  // adding a zero leaves the value of lam as it is (only the sign of a zero could differ)
      lam[g] += v;
      extra[g] = 0.0f;
    }
  }
}
