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
    if (D.flags & DET_KEEP_ALL) {
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
