// Full-tensor (9-component) material tier: update_E / update_H / reverses with 3x3 A/B matrices and
// 4-point Yee averages of the off-diagonal operands (fdtd/update.py:356-492, 610-681, 752-822,
// 932-1002; fdtd/misc.py:69-213).
//
// The reference recomputes A = M1^-1 M2, B = c M1^-1 inv per cell per step with two batched solves;
// they only depend on the materials, so the host solves once (tensor_setup.py) and the kernels read
// them (A omitted = identity when there is no conductivity).  The update needs the curl (with its
// CPML correction) at four neighbouring cells, which breaks the read-once single-sweep structure,
// so the tier runs in two phases with a (3,N) curl scratch and ping-pong field buffers:
//   phase 1  tensor_curl_kernel : K = curl(F_other) +/- CPML (+ ADE delta / c)        -> scratch
//   phase 2  tensor_apply_kernel: F'_r = sum_q A_rq avg(F_q) +/- sum_q B_rq avg(K_q)
//                                 + sources + PEC/PMC mask                             -> F_out
// One thread per cell, scalar accesses through the caches: this tier is not the throughput flagship
// (SURVEY.md section 8d counts 108 B/cell-step for it; the scratch adds 24).
#pragma once
#include "common.cuh"
#define FDTDX_BUILD_E 1
#define FDTDX_BUILD_H 1
#include "yee_kernels.cuh"  // helpers only (profiles, in_box, src_time); both kernels are compiled out
#undef FDTDX_BUILD_E
#undef FDTDX_BUILD_H

struct TensorParams {
  int nx, ny, nz;
  int wrap[3];
  float cour, dt;
  const float* F_in;     // field being updated (old values), (3,N)
  float* F_out;          // new values (ping-pong partner)
  const float* F_other;  // the field whose curl drives the update
  float* K;              // curl scratch (3,N)
  const float* A;        // (9,N) or nullptr = identity
  const float* B;        // (9,N)
  const float* mat;      // inv_eps / inv_mu as bound (tier 1|3|9) for source injection
  long long mat_cs;      // component stride of `mat` (0 for tier 1)
  int mat_tier;          // 0 scalar, 1, 3, 9
  float mat_scalar;
  const float* sc[3];    // metric scales of the curl stencil (backward for E, forward for H) or nullptr
  const float* w[3];     // cell widths for the weighted averages or nullptr
  AxisPmlDev pml[3];
  int simulate;
  int is_E;              // 1: E update (curl of H, backward differences), 0: H update
  int reverse;
  int n_walls;
  const WallDev* walls;
  int n_src;
  const SrcDev* src;
  // ADE (E only, diagonal c3; update.py:399-453)
  int n_poles;
  const float* P_cur;
  float* P_new;
  const float *c1, *c2, *c3;
  long long c_cs;
};

__device__ __forceinline__ float t_at(const TensorParams& P, const float* F, int c, int x, int y, int z) {
  if (x < 0) { if (P.wrap[0]) x += P.nx; else return 0.0f; }
  if (x >= P.nx) { if (P.wrap[0]) x -= P.nx; else return 0.0f; }
  if (y < 0) { if (P.wrap[1]) y += P.ny; else return 0.0f; }
  if (y >= P.ny) { if (P.wrap[1]) y -= P.ny; else return 0.0f; }
  if (z < 0) { if (P.wrap[2]) z += P.nz; else return 0.0f; }
  if (z >= P.nz) { if (P.wrap[2]) z -= P.nz; else return 0.0f; }
  const long long N = (long long)P.nx * P.ny * P.nz;
  return F[c * N + ((long long)x * P.ny + y) * P.nz + z];
}

__device__ __forceinline__ void t_cpml(const AxisPmlDev& A, bool is_E, int idx, long long pidx_lo, long long pidx_hi,
                                       bool simulate, float d1, float d2, float* c1, float* c2) {
  *c1 = 0.0f;
  *c2 = 0.0f;
  if (!(idx < A.lo_len || idx >= A.hi_start)) return;
  const int side = (idx >= A.hi_start) ? 1 : 0;
  const long long pidx = side ? pidx_hi : pidx_lo;
  float* q1 = is_E ? (side ? A.psiE[1][0] : A.psiE[0][0]) : (side ? A.psiH[1][0] : A.psiH[0][0]);
  float* q2 = is_E ? (side ? A.psiE[1][1] : A.psiE[0][1]) : (side ? A.psiH[1][1] : A.psiH[0][1]);
  const float a = is_E ? A.aE[idx] : A.aH[idx];
  const float b = is_E ? A.bE[idx] : A.bH[idx];
  const float km1 = is_E ? A.kE[idx] : A.kH[idx];
  float p1 = q1[pidx], p2 = q2[pidx];
  if (simulate) {
    p1 = b * p1 + a * d1;
    p2 = b * p2 + a * d2;
    q1[pidx] = p1;
    q2[pidx] = p2;
  }
  if (A.kappa_one) { *c1 = p1; *c2 = p2; }
  else { *c1 = km1 * d1 + p1; *c2 = km1 * d2 + p2; }
}

// phase 1: curl of the other field with CPML (curl.py:227-397) and, for E, the ADE polarisation
// term folded in as curl += delta / c (update.py:449-451).
__global__ void tensor_curl_kernel(const TensorParams P) {
  const long long N = (long long)P.nx * P.ny * P.nz;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= N) return;
  const int z = (int)(cell % P.nz);
  const int y = (int)((cell / P.nz) % P.ny);
  const int x = (int)(cell / ((long long)P.nz * P.ny));
  const float* G = P.F_other;
  float d[3][3];  // d[a][c] = d_a G_c (only a != c are used)
  const int s = P.is_E ? -1 : +1;
  for (int a = 0; a < 3; ++a) {
    const int ex = (a == 0) ? s : 0, ey = (a == 1) ? s : 0, ez = (a == 2) ? s : 0;
    const int idx = (a == 0) ? x : (a == 1 ? y : z);
    const float sc = P.sc[a] ? P.sc[a][idx] : 1.0f;
    for (int c = 0; c < 3; ++c) {
      if (c == a) continue;
      const float here = t_at(P, G, c, x, y, z), there = t_at(P, G, c, x + ex, y + ey, z + ez);
      float v = P.is_E ? (here - there) : (there - here);
      if (P.sc[a]) v = v * sc;
      d[a][c] = v;
    }
  }
  float K[3] = {d[1][2] - d[2][1], d[2][0] - d[0][2], d[0][1] - d[1][0]};
  const bool simulate = P.simulate && !P.reverse;
  {  // x slabs: corrects K_y (-) with d_x G_z, K_z (+) with d_x G_y
    const AxisPmlDev& A = P.pml[0];
    float c1, c2;
    t_cpml(A, P.is_E, x, ((long long)x * P.ny + y) * P.nz + z, ((long long)(x - A.hi_start) * P.ny + y) * P.nz + z, simulate, d[0][2], d[0][1], &c1, &c2);
    K[1] = K[1] - c1;
    K[2] = K[2] + c2;
  }
  {  // y slabs: corrects K_z (-) with d_y G_x, K_x (+) with d_y G_z
    const AxisPmlDev& A = P.pml[1];
    float c1, c2;
    t_cpml(A, P.is_E, y, ((long long)x * A.lo_len + y) * P.nz + z, ((long long)x * A.hi_len + (y - A.hi_start)) * P.nz + z, simulate, d[1][0], d[1][2], &c1, &c2);
    K[2] = K[2] - c1;
    K[0] = K[0] + c2;
  }
  {  // z slabs: corrects K_x (-) with d_z G_y, K_y (+) with d_z G_x
    const AxisPmlDev& A = P.pml[2];
    float c1, c2;
    t_cpml(A, P.is_E, z, ((long long)x * P.ny + y) * A.lo_len + z, ((long long)x * P.ny + y) * A.hi_len + (z - A.hi_start), simulate, d[2][1], d[2][0], &c1, &c2);
    K[0] = K[0] - c1;
    K[1] = K[1] + c2;
  }
  if (P.is_E && P.n_poles > 0 && !P.reverse) {
    for (int c = 0; c < 3; ++c) {
      const float Eo = P.F_in[c * N + cell];
      float delta = 0.0f;
      for (int p = 0; p < P.n_poles; ++p) {
        const long long pi = (long long)p * 3 * N + c * N + cell;
        const long long ci = (long long)p * (P.c_cs ? 3 * N : N) + c * P.c_cs + cell;
        const float Pc = P.P_cur[pi], Pp = P.P_new[pi];
        const float Phat = (P.c1[ci] * Pc + P.c2[ci] * Pp) + P.c3[ci] * Eo;
        const float dd = Pc - Phat;
        delta = (p == 0) ? dd : delta + dd;
        P.P_new[pi] = Phat;
      }
      K[c] = K[c] + delta / P.cour;
    }
  }
  P.K[cell] = K[0];
  P.K[N + cell] = K[1];
  P.K[2 * N + cell] = K[2];
}

// 4-point average of component c of array F at the Yee location of component l (fdtd/misc.py:132-213)
__device__ __forceinline__ float t_avg(const TensorParams& P, const float* F, int c, int l, int x, int y, int z) {
  const int lx = (l == 0), ly = (l == 1), lz = (l == 2);
  const int cx = (c == 0), cy = (c == 1), cz = (c == 2);
  if (P.is_E) {
    const float s00 = t_at(P, F, c, x, y, z);
    const float s10 = t_at(P, F, c, x + lx, y + ly, z + lz);
    const float s01 = t_at(P, F, c, x - cx, y - cy, z - cz);
    const float s11 = t_at(P, F, c, x + lx - cx, y + ly - cy, z + lz - cz);
    if (P.w[0] == nullptr) return (((s00 + s10) + s01) + s11) / 4.0f;
    const int i = (c == 0) ? x : (c == 1 ? y : z);
    const float width = P.w[c][i], prev = P.w[c][i > 0 ? i - 1 : 0];
    const float cen = 0.5f * (s00 + s10), cen_m = 0.5f * (s01 + s11);
    return (cen * prev + cen_m * width) / (width + prev);
  } else {
    const float s00 = t_at(P, F, c, x, y, z);
    const float s10 = t_at(P, F, c, x - lx, y - ly, z - lz);
    const float s01 = t_at(P, F, c, x + cx, y + cy, z + cz);
    const float s11 = t_at(P, F, c, x - lx + cx, y - ly + cy, z - lz + cz);
    if (P.w[0] == nullptr) return (((s00 + s10) + s01) + s11) / 4.0f;
    const int i = (l == 0) ? x : (l == 1 ? y : z);
    const float width = P.w[l][i], prev = P.w[l][i > 0 ? i - 1 : 0];
    const float e0 = (s00 * prev + s10 * width) / (width + prev);
    const float e1 = (s01 * prev + s11 * width) / (width + prev);
    return 0.5f * (e0 + e1);
  }
}

__device__ __forceinline__ float t_mat(const TensorParams& P, int row, int col, long long cell) {
  if (P.mat_tier == 9) return P.mat[(long long)(row * 3 + col) * P.mat_cs + cell];
  if (row != col) return 0.0f;
  if (P.mat_tier == 0) return P.mat_scalar;
  return P.mat[(long long)(P.mat_tier == 1 ? 0 : row) * P.mat_cs + cell];
}

// source injection at one cell for any material tier (tfsf.py:285-307, 384-408; dipole.py:207-232)
static __device__ __noinline__ void t_inject(const TensorParams& P, int t, bool inverse, int x, int y, int z, long long cell, float* Fv) {
  for (int s = 0; s < P.n_src; ++s) {
    const SrcDev& S = P.src[s];
    if (!in_box(S.lo, S.hi, x, y, z)) continue;
    float tf;
    if (!src_time(S, t, P.is_E ? 0.0f : 0.5f, &tf)) continue;
    if (S.kind == 0) {
      const int fy = S.hi[1] - S.lo[1], fz = S.hi[2] - S.lo[2];
      const long long fn = (long long)(S.hi[0] - S.lo[0]) * fy * fz;
      const long long f = ((long long)(x - S.lo[0]) * fy + (y - S.lo[1])) * fz + (z - S.lo[2]);
      const int n = S.normal_axis, a = (n + 1) % 3, b = (n + 2) % 3;
      const float sign = inverse ? -S.sign : S.sign;
      const float* toff = P.is_E ? S.toffH : S.toffE;
      const float* inc = P.is_E ? S.Hinc : S.Einc;
      const float c = P.is_E ? S.cE : S.cH;
      const float amp_a = src_profile(S, (tf + toff[a * fn + f]) * P.dt) * S.static_amp;
      const float amp_b = src_profile(S, (tf + toff[b * fn + f]) * P.dt) * S.static_amp;
      float Ia = inc[a * fn + f] * amp_a, Ib = inc[b * fn + f] * amp_b;
      const float* incI = P.is_E ? S.HincI : S.EincI;
      if (incI != nullptr) {
        Ia = Ia + incI[a * fn + f] * (src_profile_ph(S, (tf + toff[a * fn + f]) * P.dt, S.pq) * S.static_amp);
        Ib = Ib + incI[b * fn + f] * (src_profile_ph(S, (tf + toff[b * fn + f]) * P.dt, S.pq) * S.static_amp);
      }
      if (P.mat_tier == 9) {
        const int rows[3] = {n, a, b};
        for (int q = 0; q < 3; ++q) {
          const int row = rows[q];
          float corr;
          if (P.is_E) corr = c * (t_mat(P, row, a, cell) * (+Ib) + t_mat(P, row, b, cell) * (-Ia));
          else corr = c * (t_mat(P, row, a, cell) * (-Ib) + t_mat(P, row, b, cell) * (+Ia));
          Fv[row] = Fv[row] + sign * corr;
        }
      } else if (P.is_E) {
        const float Hb = (Ib * c) * t_mat(P, a, a, cell), Ha = (Ia * c) * t_mat(P, b, b, cell);
        Fv[a] = Fv[a] + sign * Hb;
        Fv[b] = Fv[b] + (-sign) * Ha;
      } else {
        const float Ea = (Ia * c) * t_mat(P, b, b, cell), Eb = (Ib * c) * t_mat(P, a, a, cell);
        Fv[b] = Fv[b] + sign * Ea;
        Fv[a] = Fv[a] + (-sign) * Eb;
      }
    } else if ((S.electric != 0) == (P.is_E != 0)) {
      const float amp = src_profile(S, tf * P.dt);
      const float sg = inverse ? 1.0f : -1.0f;
      const float scale = S.dip_scale * amp;
      if (P.mat_tier == 9) {
        for (int axis = 0; axis < 3; ++axis) Fv[axis] = Fv[axis] + sg * (scale * t_mat(P, axis, S.pol, cell));
      } else {
        Fv[S.pol] = Fv[S.pol] + sg * (scale * t_mat(P, S.pol, S.pol, cell));
      }
    }
  }
}

// reverse pass: sources are removed from the field *before* the averages are taken
// (update.py:558-584 / 877-903); in place on F_in's buffer (passed as F_out here).
__global__ void tensor_inject_kernel(const TensorParams P, const int t, const int src_index, const int inverse = 1) {
  const SrcDev& S = P.src[src_index];
  const int ex = S.hi[0] - S.lo[0], ey = S.hi[1] - S.lo[1], ez = S.hi[2] - S.lo[2];
  const long long n = (long long)ex * ey * ez;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int z = S.lo[2] + (int)(idx % ez), y = S.lo[1] + (int)((idx / ez) % ey), x = S.lo[0] + (int)(idx / ((long long)ez * ey));
  if (x < 0 || x >= P.nx) return;
  const long long N = (long long)P.nx * P.ny * P.nz;
  const long long cell = ((long long)x * P.ny + y) * P.nz + z;
  float Fv[3] = {P.F_out[cell], P.F_out[N + cell], P.F_out[2 * N + cell]};
  TensorParams R = P;  // apply only source `src_index` (sources are independent additive terms)
  R.src = P.src + src_index;
  R.n_src = 1;
  t_inject(R, t, inverse != 0, x, y, z, cell, Fv);
  if (!inverse) {  // forward order: update -> sources -> PEC / PMC mask
    for (int w = 0; w < P.n_walls; ++w) {
      const WallDev W = P.walls[w];
      if (W.kind == (P.is_E ? 0 : 1) && in_box(W.lo, W.hi, x, y, z)) {
        if (W.axis != 0) Fv[0] = 0.0f;
        if (W.axis != 1) Fv[1] = 0.0f;
        if (W.axis != 2) Fv[2] = 0.0f;
      }
    }
  }
  P.F_out[cell] = Fv[0];
  P.F_out[N + cell] = Fv[1];
  P.F_out[2 * N + cell] = Fv[2];
}

// phase 2
__global__ void tensor_apply_kernel(const TensorParams P, const int t) {
  const long long N = (long long)P.nx * P.ny * P.nz;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= N) return;
  const int z = (int)(cell % P.nz);
  const int y = (int)((cell / P.nz) % P.ny);
  const int x = (int)(cell / ((long long)P.nz * P.ny));
  // forward E: A F + B K ; forward H: A F - B K ; reverse E: A F - B K ; reverse H: A F + B K
  const bool plus = (P.is_E != 0) != (P.reverse != 0);
  float out[3];
  for (int r = 0; r < 3; ++r) {
    float fa[3], ka[3];
    for (int q = 0; q < 3; ++q) {
      fa[q] = (q == r) ? P.F_in[q * N + cell] : t_avg(P, P.F_in, q, r, x, y, z);
      ka[q] = (q == r) ? P.K[q * N + cell] : t_avg(P, P.K, q, r, x, y, z);
    }
    float t1;
    if (P.A) t1 = (P.A[(long long)(3 * r + 0) * N + cell] * fa[0] + P.A[(long long)(3 * r + 1) * N + cell] * fa[1]) + P.A[(long long)(3 * r + 2) * N + cell] * fa[2];
    else t1 = fa[r];
    const float t2 = (P.B[(long long)(3 * r + 0) * N + cell] * ka[0] + P.B[(long long)(3 * r + 1) * N + cell] * ka[1]) + P.B[(long long)(3 * r + 2) * N + cell] * ka[2];
    out[r] = plus ? (t1 + t2) : (t1 - t2);
  }
  if (!P.reverse && P.n_src > 0) t_inject(P, t, false, x, y, z, cell, out);
  for (int w = 0; w < P.n_walls; ++w) {
    const WallDev W = P.walls[w];
    if (W.kind == (P.is_E ? 0 : 1) && in_box(W.lo, W.hi, x, y, z)) {
      if (W.axis != 0) out[0] = 0.0f;
      if (W.axis != 1) out[1] = 0.0f;
      if (W.axis != 2) out[2] = 0.0f;
    }
  }
  P.F_out[cell] = out[0];
  P.F_out[N + cell] = out[1];
  P.F_out[2 * N + cell] = out[2];
}

// phase 2, four cells per thread with 128-bit accesses (uniform grid, Nz % 4 == 0, aligned buffers):
// same samples and the same summation order as tensor_apply_kernel / t_avg.  Sources are applied
// afterwards by tensor_inject_kernel (O(surface)); blockDim (32, 8), grid (z tiles, y tiles, x planes).
// HAS_A = false: no conductivity, A is the identity (fdtd/misc.py:88-96) and the field averages are not needed
template <bool IS_E, bool HAS_A>
__global__ void __launch_bounds__(256) tensor_apply4_kernel(const TensorParams P) {
  constexpr int V = 4;
  const int k0 = (blockIdx.x * 32 + threadIdx.x) * V;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  if (k0 >= P.nz || y >= P.ny) return;
  const long long plane = (long long)P.ny * P.nz, N = plane * P.nx;
  const long long cell0 = (long long)x * plane + (long long)y * P.nz + k0;
  const bool plus = IS_E != (P.reverse != 0);
  // component c of F at cells (x+dx, y+dy, k0+dz .. k0+dz+3), halo rule of pad_fields
  auto ldsh = [&](const float* F, int c, int dx, int dy, int dz) -> Vec<V> {
    int xx = x + dx, yy = y + dy;
    if (xx < 0) { if (P.wrap[0]) xx += P.nx; else return zerov<V>(); }
    if (xx >= P.nx) { if (P.wrap[0]) xx -= P.nx; else return zerov<V>(); }
    if (yy < 0) { if (P.wrap[1]) yy += P.ny; else return zerov<V>(); }
    if (yy >= P.ny) { if (P.wrap[1]) yy -= P.ny; else return zerov<V>(); }
    const float* row = F + c * N + (long long)xx * plane + (long long)yy * P.nz;
    const Vec<V> v = ldv<V>(row + k0);
    if (dz == 0) return v;
    int kz = (dz > 0) ? k0 + V : k0 - 1;
    float edge = 0.0f;
    bool ok = true;
    if (kz < 0) { if (P.wrap[2]) kz = P.nz - 1; else ok = false; }
    if (kz >= P.nz) { if (P.wrap[2]) kz = 0; else ok = false; }
    if (ok) edge = row[kz];
    Vec<V> r;
    if (dz > 0) { r.v[0] = v.v[1]; r.v[1] = v.v[2]; r.v[2] = v.v[3]; r.v[3] = edge; }
    else { r.v[0] = edge; r.v[1] = v.v[0]; r.v[2] = v.v[1]; r.v[3] = v.v[2]; }
    return r;
  };
  Vec<V> F0[3], K0[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) { F0[q] = ldv<V>(P.F_in + q * N + cell0); K0[q] = ldv<V>(P.K + q * N + cell0); }
  Vec<V> out[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    Vec<V> fa[3], ka[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (q == r) { fa[q] = F0[q]; ka[q] = K0[q]; continue; }
      // t_avg(F, c = q, l = r): E: p, p + e_l, p - e_c, p + e_l - e_c ; H: p, p - e_l, p + e_c, p - e_l + e_c
      const int sl = IS_E ? +1 : -1, scn = IS_E ? -1 : +1;
      const int lx = (r == 0) * sl, ly = (r == 1) * sl, lz = (r == 2) * sl;
      const int cx = (q == 0) * scn, cy = (q == 1) * scn, cz = (q == 2) * scn;
      const Vec<V> k10 = ldsh(P.K, q, lx, ly, lz), k01 = ldsh(P.K, q, cx, cy, cz), k11 = ldsh(P.K, q, lx + cx, ly + cy, lz + cz);
#pragma unroll
      for (int e = 0; e < V; ++e) ka[q].v[e] = (((K0[q].v[e] + k10.v[e]) + k01.v[e]) + k11.v[e]) / 4.0f;
      if (HAS_A) {
        const Vec<V> f10 = ldsh(P.F_in, q, lx, ly, lz), f01 = ldsh(P.F_in, q, cx, cy, cz), f11 = ldsh(P.F_in, q, lx + cx, ly + cy, lz + cz);
#pragma unroll
        for (int e = 0; e < V; ++e) fa[q].v[e] = (((F0[q].v[e] + f10.v[e]) + f01.v[e]) + f11.v[e]) / 4.0f;
      }
    }
    Vec<V> b0 = ldv<V>(P.B + (long long)(3 * r + 0) * N + cell0), b1 = ldv<V>(P.B + (long long)(3 * r + 1) * N + cell0), b2 = ldv<V>(P.B + (long long)(3 * r + 2) * N + cell0);
    Vec<V> a0, a1, a2;
    if (HAS_A) { a0 = ldv<V>(P.A + (long long)(3 * r + 0) * N + cell0); a1 = ldv<V>(P.A + (long long)(3 * r + 1) * N + cell0); a2 = ldv<V>(P.A + (long long)(3 * r + 2) * N + cell0); }
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float t1 = HAS_A ? (a0.v[e] * fa[0].v[e] + a1.v[e] * fa[1].v[e]) + a2.v[e] * fa[2].v[e] : fa[r].v[e];
      const float t2 = (b0.v[e] * ka[0].v[e] + b1.v[e] * ka[1].v[e]) + b2.v[e] * ka[2].v[e];
      out[r].v[e] = plus ? (t1 + t2) : (t1 - t2);
    }
  }
  for (int w = 0; w < P.n_walls; ++w) {
    const WallDev W = P.walls[w];
    if (W.kind != (IS_E ? 0 : 1) || x < W.lo[0] || x >= W.hi[0] || y < W.lo[1] || y >= W.hi[1]) continue;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      if (k0 + e >= W.lo[2] && k0 + e < W.hi[2]) {
        if (W.axis != 0) out[0].v[e] = 0.0f;
        if (W.axis != 1) out[1].v[e] = 0.0f;
        if (W.axis != 2) out[2].v[e] = 0.0f;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) stv<V>(P.F_out + r * N + cell0, out[r]);
}
