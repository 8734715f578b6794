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

__global__ void adj_local_kernel(const AdjParams P) {
  const long long N = (long long)P.nx * P.ny * P.nz;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= N) return;
  const int z = (int)(cell % P.nz);
  const int y = (int)((cell / P.nz) % P.ny);
  const int x = (int)(cell / ((long long)P.nz * P.ny));
  const int pos[3] = {x, y, z};
  // ---- primal derivatives and curl at this cell (same arithmetic as the forward kernels) ----
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
  float K[3] = {d[1][2] - d[2][1], d[2][0] - d[0][2], d[0][1] - d[1][0]};
  // CPML state of this cell per axis
  bool inp[3];
  float ca[3], cb[3], ck[3];
  long long pidx[3];
  int side[3];
  for (int a = 0; a < 3; ++a) {
    const AxisPmlDev& A = P.pml[a];
    const int idx = pos[a];
    inp[a] = (idx < A.lo_len || idx >= A.hi_start);
    side[a] = (idx >= A.hi_start) ? 1 : 0;
    ca[a] = cb[a] = ck[a] = 0.0f;
    pidx[a] = 0;
    if (!inp[a]) continue;
    ca[a] = P.is_E ? A.aE[idx] : A.aH[idx];
    cb[a] = P.is_E ? A.bE[idx] : A.bH[idx];
    ck[a] = A.kappa_one ? 0.0f : (P.is_E ? A.kE[idx] : A.kH[idx]);
    if (a == 0) pidx[a] = ((long long)(side[a] ? x - A.hi_start : x) * P.ny + y) * P.nz + z;
    else if (a == 1) pidx[a] = ((long long)x * (side[a] ? A.hi_len : A.lo_len) + (side[a] ? y - A.hi_start : y)) * P.nz + z;
    else pidx[a] = ((long long)x * P.ny + y) * (side[a] ? A.hi_len : A.lo_len) + (side[a] ? z - A.hi_start : z);
    const int i = (a + 1) % 3, j = (a + 2) % 3;
    const float* q1 = P.is_E ? (side[a] ? A.psiE[1][0] : A.psiE[0][0]) : (side[a] ? A.psiH[1][0] : A.psiH[0][0]);
    const float* q2 = P.is_E ? (side[a] ? A.psiE[1][1] : A.psiE[0][1]) : (side[a] ? A.psiH[1][1] : A.psiH[0][1]);
    const float p1 = cb[a] * q1[pidx[a]] + ca[a] * d[a][j];
    const float p2 = cb[a] * q2[pidx[a]] + ca[a] * d[a][i];
    K[i] = K[i] - (ck[a] * d[a][j] + p1);
    K[j] = K[j] + (ck[a] * d[a][i] + p2);
  }
  // ---- transpose of the material update ----
  float lam[3] = {P.lamF[cell], P.lamF[N + cell], P.lamF[2 * N + cell]};
  for (int w = 0; w < P.n_walls; ++w) {
    const WallDev W = P.walls[w];
    if (W.kind == (P.is_E ? 0 : 1) && x >= W.lo[0] && x < W.hi[0] && y >= W.lo[1] && y < W.hi[1] && z >= W.lo[2] && z < W.hi[2]) {
      if (W.axis != 0) lam[0] = 0.0f;
      if (W.axis != 1) lam[1] = 0.0f;
      if (W.axis != 2) lam[2] = 0.0f;
    }
  }
  float lamK[3];
  float gacc = 0.0f;
  for (int c = 0; c < 3; ++c) {
    const float m = P.mat_tier == 0 ? P.mat_scalar : P.mat[(long long)(P.mat_tier == 1 ? 0 : c) * P.mat_cs + cell];
    float alpha = 0.0f;
    if (P.sig) {
      const float sg = P.sig[c * P.sig_cs + cell];
      alpha = P.is_E ? ((P.cour * sg) * P.eta0) / 2.0f : ((P.cour * sg) / P.eta0) / 2.0f;
    }
    const float sv = alpha * m;
    const float u = lam[c] / (1.0f + sv);
    const float Fc = P.F[c * N + cell];
    const float cK = P.cour * K[c];
    const float Fpre = P.is_E ? ((1.0f - sv) * Fc + cK * m) / (1.0f + sv) : ((1.0f - sv) * Fc - cK * m) / (1.0f + sv);
    const float g = P.is_E ? u * (cK - alpha * Fc - alpha * Fpre) : u * (-cK - alpha * Fc - alpha * Fpre);
    if (P.g_mat) {
      if (P.mat_tier == 3) P.g_mat[c * N + cell] += g;
      else gacc += g;
    }
    float lin = (1.0f - sv) * u;
    if (P.lam_extra) lin += P.lam_extra[c * N + cell];
    P.lamF[c * N + cell] = lin;
    lamK[c] = P.is_E ? (P.cour * m) * u : -(P.cour * m) * u;
  }
  if (P.g_mat && P.mat_tier == 1) P.g_mat[cell] += gacc;
  // ---- transpose of curl + CPML: cotangent of each derivative d_a G_c ----
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
    P.ld[(long long)(2 * a + 0) * N + cell] = l1;
    P.ld[(long long)(2 * a + 1) * N + cell] = l2;
  }
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

// ---------------------------------------------------------------------------------------------
// detector cotangents
// ---------------------------------------------------------------------------------------------
struct DetAdj {
  const float* cot[4];  // cotangent of the detector state (same layout as the state)
  float* lamE;          // (3,N) cotangent of E'
  float* lamH;          // (3,N) cotangent of H'
  float* lamHprev;      // (3,N) cotangent of the step's input H (through H_bar = (H + H')/2)
  float* g_eps;         // inv_eps gradient (energy detectors) or nullptr
  int eps_tier;
};

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
        g = A.cot[0][((long long)slot * ex + rx) * ey + ry] / (float)ez + A.cot[1][((long long)slot * ex + rx) * ez + rz] / (float)ey +
            A.cot[2][((long long)slot * ey + ry) * ez + rz] / (float)ex;
      else
        g = (rz == D.slice_idx[2] ? A.cot[0][((long long)slot * ex + rx) * ey + ry] : 0.0f) +
            (ry == D.slice_idx[1] ? A.cot[1][((long long)slot * ex + rx) * ez + rz] : 0.0f) +
            (rx == D.slice_idx[0] ? A.cot[2][((long long)slot * ey + ry) * ez + rz] : 0.0f);
    } else g = A.cot[0][(long long)slot * n + cell];
    const long long N = (long long)G.nx * G.ny * G.nz;
    const long long gidx = ((long long)x * G.ny + y) * G.nz + z;
    for (int c = 0; c < 3; ++c) {
      const float ie = G.eps[c * G.eps_cs + gidx];
      const float im = G.mu ? G.mu[c * G.mu_cs + gidx] : G.inv_mu_scalar;
      lE[c] = g * Es[c] / ie;
      lH[c] = g * Hs[c] / im;
      if (A.g_eps) atomicAdd(A.g_eps + (A.eps_tier == 1 ? 0 : c) * N + gidx, g * (-0.5f * Es[c] * Es[c] / (ie * ie)));
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
