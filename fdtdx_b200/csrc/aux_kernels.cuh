// Detector accumulation, PML-interface record/replay and field-reset kernels (sm_100a).
//
//   det_*      : update_detector_states + interpolate_fields + the per-type Detector.update
//                (fdtd/update.py:1040-1137, core/physics/curl.py:42-224, objects/detectors/*.py,
//                 core/physics/metrics.py:15-117)
//   rec_*      : collect_interfaces / add_interfaces + Recorder.compress/decompress
//                (fdtd/update.py:1140-1222, fdtd/misc.py:10-66, interfaces/recorder.py:70-199,
//                 interfaces/modules.py:138-161, interfaces/time_filter.py:237-250)
//   reset_pml  : PerfectlyMatchedLayer.apply_field_reset (perfectly_matched_layer.py:226-229)
//
// These touch O(surface) data (or a detector's own region), so they are separate small launches
// after the H half-step rather than an epilogue of it: the co-location stencil needs H_new at
// neighbouring cells, which only exist once the whole H kernel has finished.
#pragma once
#include "common.cuh"

struct GridDev {
  int nx, ny, nz;
  int x_offset;  // global x of local plane 0 (x-slab sharding)
  int wrap[3];
  const float* E;
  const float* H;
  const float* eps;
  const float* mu;
  long long eps_cs, mu_cs;
  int eps_tier, mu_tier;  // 1 | 3 | 9 (mu: 0 = scalar)
  float inv_mu_scalar;
  const float* w[3];  // cell widths per axis (global indexing for x) or nullptr on a uniform grid
  // Bloch axes with k != 0 (two real systems, see StepParams::bH): partner arrays and (cos, +-sin)
  const float* Ep;
  const float* Hp;
  float bc[3], bs[3];
  // x-slab sharding: plane x0-1 of the lower neighbour rank, (3,ny,nz) each, for detector stencils at local
  // x = 0 (E, H of this step; H before this step's H update) - nullptr: halo rule (zero / wrap)
  const float* xlo_E;
  const float* xlo_H;
  const float* xlo_Hp;
  // config.symmetry: sym[a] - the min-side halo of axis a is never wrapped (update.py:121-125); mirror[a] - an
  // electric symmetry wall sits on the min face, and the detector stencil reads the parity-weighted mirror
  // partner there instead of the zero halo (pad_fields_with_symmetry_mirror, update.py:139-198)
  int sym[3], mirror[3];
};

struct DetDev {
  int kind, flags;
  int lo[3], hi[3];
  int comp_mask, ncomp, aux;
  const uint8_t* on;
  const int32_t* arr_idx;
  const float* weights;  // region-shaped (or (3,region) for keep_all Poynting) or nullptr
  float wsum;
  int nf;
  const float2* ph_table;  // (T, nf)
  const float* window;     // (T)
  float scale;
  int slice_idx[3];
  float* hprev;    // (3, ex+1, ey+1, ez+1): H before this step's H update, halo rules applied
                   // DET_VOLUME: (3, ex+1, ey+1, hrow) with 16-byte aligned rows, see det_volume.cuh
  float* scratch;  // (nvals, region) staging for reductions
  int hrow, hz0;   // DET_VOLUME: row length of hprev and the global z of its element 0 (a multiple of 4)
  float* part[3];  // DET_VOLUME slice means: partial sums over z tiles / y tiles / x chunks
  int npart[3];
  float* hprev_full;  // DET_VOLUME: plan-wide (3,Nx,Ny,Nz) copy of H before this step's H update (det_volume.cuh)
  float* state[4];
};

#define DET_EXACT 1
#define DET_INVERSE 2
#define DET_REDUCE 4
#define DET_SLICES 8
#define DET_SLICE_MEAN 16
#define DET_KEEP_ALL 32
#define DET_NEGATIVE 64
#define DET_CLOSED 256  // Poynting: net flux through the box faces (aux = bit mask of the active axes)
#define DET_VOLUME 128  // large exact-interpolation region: row-marching gather / sample kernels (det_volume.cuh)

// CHK = false: the caller guarantees that (x, y, z) is inside the local grid (interior fast path of the
// detector stencil: no halo rule, no wrap)
template <bool CHK>
__device__ __forceinline__ float grid_at_t(const GridDev& G, const float* F, int c, int x, int y, int z);
__device__ __forceinline__ float grid_at(const GridDev& G, const float* F, int c, int x, int y, int z) {
  int side[3] = {0, 0, 0};  // -1: low-side ghost (x conj(phase)), +1: high-side ghost (x phase)
  if (G.sym[0] | G.sym[1] | G.sym[2]) {
    // symmetric axes: index -1 maps per axis to its mirror partner (electric wall: parity * the first cell for a
    // component sampled half a cell off the plane, the second for one sampled on it) or to the zero halo
    const bool isE = (F == G.E);
    int p[3] = {x, y, z};
    const int n[3] = {G.nx, G.ny, G.nz};
    float par = 1.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (p[a] < 0 && G.sym[a]) {
        if (!G.mirror[a] || !(isE || F == G.H)) return 0.0f;
        const bool normal = (c == a);
        const bool on_plane = isE ? !normal : normal;      // component_sits_on_plane
        par = (isE ? normal : !normal) ? par : -par;        // field_component_parity, wall = -1
        p[a] = on_plane ? 1 : 0;
        if (p[a] >= n[a]) return 0.0f;
      }
    }
    if (p[0] != x || p[1] != y || p[2] != z) {
      x = p[0]; y = p[1]; z = p[2];
      // the remaining out-of-range coordinates follow the ordinary halo rule
      if (x < 0) { if (G.wrap[0]) x += G.nx; else return 0.0f; }
      if (x >= G.nx) { if (G.wrap[0]) x -= G.nx; else return 0.0f; }
      if (y < 0) { if (G.wrap[1]) y += G.ny; else return 0.0f; }
      if (y >= G.ny) { if (G.wrap[1]) y -= G.ny; else return 0.0f; }
      if (z < 0) { if (G.wrap[2]) z += G.nz; else return 0.0f; }
      if (z >= G.nz) { if (G.wrap[2]) z -= G.nz; else return 0.0f; }
      const long long Ns = (long long)G.nx * G.ny * G.nz;
      return par * F[c * Ns + ((long long)x * G.ny + y) * G.nz + z];
    }
  }
  if (x < 0) {
    const float* X = (F == G.E) ? G.xlo_E : ((F == G.H) ? G.xlo_H : nullptr);
    if (X != nullptr) {  // the lower neighbour rank's last plane
      if (y < 0) { if (G.wrap[1]) y += G.ny; else return 0.0f; }
      if (y >= G.ny) { if (G.wrap[1]) y -= G.ny; else return 0.0f; }
      if (z < 0) { if (G.wrap[2]) z += G.nz; else return 0.0f; }
      if (z >= G.nz) { if (G.wrap[2]) z -= G.nz; else return 0.0f; }
      return X[((long long)c * G.ny + y) * G.nz + z];
    }
    if (G.wrap[0]) { x += G.nx; side[0] = -1; } else return 0.0f;
  }
  if (x >= G.nx) { if (G.wrap[0]) { x -= G.nx; side[0] = 1; } else return 0.0f; }
  if (y < 0) { if (G.wrap[1]) { y += G.ny; side[1] = -1; } else return 0.0f; }
  if (y >= G.ny) { if (G.wrap[1]) { y -= G.ny; side[1] = 1; } else return 0.0f; }
  if (z < 0) { if (G.wrap[2]) { z += G.nz; side[2] = -1; } else return 0.0f; }
  if (z >= G.nz) { if (G.wrap[2]) { z -= G.nz; side[2] = 1; } else return 0.0f; }
  const long long N = (long long)G.nx * G.ny * G.nz;
  const long long idx = c * N + ((long long)x * G.ny + y) * G.nz + z;
  float a = F[idx];
  if (G.Ep != nullptr && (side[0] | side[1] | side[2])) {
    // Bloch ghost: the pad corrections of the wrapped axes are applied one after the other (bloch.py:61-96)
    const float* Fp = (F == G.E) ? G.Ep : ((F == G.H) ? G.Hp : nullptr);
    if (Fp != nullptr) {
      float b = Fp[idx];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        if (side[ax] == 0) continue;
        const float cc = G.bc[ax], ss = (side[ax] < 0) ? G.bs[ax] : -G.bs[ax];
        const float na = a * cc + b * ss, nb = b * cc - a * ss;
        a = na; b = nb;
      }
    }
  }
  return a;
}

// _backward_edge_average (curl.py:42-83)
__device__ __forceinline__ float bea(const GridDev& G, float cur, float prev, int axis, int idx) {
  const float* w = G.w[axis];
  if (w == nullptr) return 0.5f * (cur + prev);
  const int gi = idx + (axis == 0 ? G.x_offset : 0);
  const float chw = 0.5f * w[gi];
  const float phw = 0.5f * w[gi > 0 ? gi - 1 : 0];
  return (cur * phw + prev * chw) / (chw + phw);
}

// Copy H (region + one-cell halo on the low x/y side, +1 on the high z side) before the H update.
__device__ __forceinline__ void det_gather_body(const GridDev& G, const DetDev& D);
__global__ void det_gather_hprev_kernel(const GridDev G, const DetDev D) { det_gather_body(G, D); }
// batched: one launch for all detectors (blockIdx.y = detector), gated on the device by on[t]
// (the descriptor is staged in shared memory: read through a global reference it is re-loaded after
// every store, because the stores may alias it)
__device__ __forceinline__ void det_stage_descriptor(DetDev* dst, const DetDev* src) {
  pdl_trigger();
  for (int q = threadIdx.x; q < (int)(sizeof(DetDev) / 4); q += blockDim.x) reinterpret_cast<int*>(dst)[q] = reinterpret_cast<const int*>(src)[q];
  __syncthreads();
  pdl_wait();
}
__global__ void det_gather_batch_kernel(const GridDev G, const DetDev* __restrict__ dets, const int t, const int inverse) {
  __shared__ DetDev sD;
  det_stage_descriptor(&sD, dets + blockIdx.y);
  const DetDev& D = sD;
  if (((D.flags & DET_INVERSE) != 0) != (inverse != 0) || !D.on[t] || !(D.flags & DET_EXACT) || (D.flags & DET_VOLUME)) return;
  det_gather_body(G, D);
}
__device__ __forceinline__ void det_gather_body(const GridDev& G, const DetDev& D) {
  const int sx = D.hi[0] - D.lo[0] + 1, sy = D.hi[1] - D.lo[1] + 1, sz = D.hi[2] - D.lo[2] + 1;
  const long long n = (long long)sx * sy * sz;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < 3 * n; idx += (long long)gridDim.x * blockDim.x) {
    // 32-bit index arithmetic whenever the region allows it (64-bit div/mod costs ~100 instructions each)
    int c, a, b, g;
    if (3 * n < 0x7fffffffLL) {
      const unsigned u = (unsigned)idx, un = (unsigned)n;
      c = (int)(u / un);
      unsigned r = u - (unsigned)c * un;
      g = (int)(r % (unsigned)sz); r /= (unsigned)sz;
      b = (int)(r % (unsigned)sy);
      a = (int)(r / (unsigned)sy);
    } else {
      c = (int)(idx / n);
      long long r = idx - c * n;
      g = (int)(r % sz); r /= sz;
      b = (int)(r % sy);
      a = (int)(r / sy);
    }
    // plane slices without averaging only sample three planes: copy just the H cells their stencils read
    if ((D.flags & DET_SLICES) && !(D.flags & DET_SLICE_MEAN) && !(a == D.slice_idx[0] || a == D.slice_idx[0] + 1 || b == D.slice_idx[1] ||
                                                                    b == D.slice_idx[1] + 1 || g == D.slice_idx[2] || g == D.slice_idx[2] + 1))
      continue;
    D.hprev[idx] = grid_at(G, G.H, c, D.lo[0] - 1 + a, D.lo[1] - 1 + b, D.lo[2] + g);
  }
}

template <bool CHK>
__device__ __forceinline__ float grid_at_t(const GridDev& G, const float* F, int c, int x, int y, int z) {
  if (CHK) return grid_at(G, F, c, x, y, z);
  const long long N = (long long)G.nx * G.ny * G.nz;
  return F[c * N + ((long long)x * G.ny + y) * G.nz + z];
}
template <bool CHK = true>
__device__ __forceinline__ float hbar(const GridDev& G, const DetDev& D, int c, int x, int y, int z) {
  const int sy = D.hi[1] - D.lo[1] + 1, sz = D.hi[2] - D.lo[2] + 1, sx = D.hi[0] - D.lo[0] + 1;
  const long long n = (long long)sx * sy * sz;
  const float hp = D.hprev[c * n + ((long long)(x - D.lo[0] + 1) * sy + (y - D.lo[1] + 1)) * sz + (z - D.lo[2])];
  return (hp + grid_at_t<CHK>(G, G.H, c, x, y, z)) / 2.0f;
}

// interpolate_fields at one cell (curl.py:86-224): everything co-located at the Ez point.
template <bool CHK>
__device__ __forceinline__ void colocate_t(const GridDev& G, const DetDev& D, int x, int y, int z, float* Es, float* Hs) {
  if (!(D.flags & DET_EXACT)) {
    for (int c = 0; c < 3; ++c) {
      Es[c] = grid_at_t<CHK>(G, G.E, c, x, y, z);
      Hs[c] = grid_at_t<CHK>(G, G.H, c, x, y, z);
    }
    return;
  }
  const float* E = G.E;
  float lo = bea(G, grid_at_t<CHK>(G, E, 0, x, y, z), grid_at_t<CHK>(G, E, 0, x - 1, y, z), 0, x);
  float hi = bea(G, grid_at_t<CHK>(G, E, 0, x, y, z + 1), grid_at_t<CHK>(G, E, 0, x - 1, y, z + 1), 0, x);
  Es[0] = (lo + hi) / 2.0f;
  lo = bea(G, grid_at_t<CHK>(G, E, 1, x, y, z), grid_at_t<CHK>(G, E, 1, x, y - 1, z), 1, y);
  hi = bea(G, grid_at_t<CHK>(G, E, 1, x, y, z + 1), grid_at_t<CHK>(G, E, 1, x, y - 1, z + 1), 1, y);
  Es[1] = (lo + hi) / 2.0f;
  Es[2] = grid_at_t<CHK>(G, E, 2, x, y, z);
  Hs[0] = bea(G, hbar<CHK>(G, D, 0, x, y, z), hbar<CHK>(G, D, 0, x, y - 1, z), 1, y);
  Hs[1] = bea(G, hbar<CHK>(G, D, 1, x, y, z), hbar<CHK>(G, D, 1, x - 1, y, z), 0, x);
  float lx = bea(G, hbar<CHK>(G, D, 2, x, y, z), hbar<CHK>(G, D, 2, x - 1, y, z), 0, x);
  float lxm = bea(G, hbar<CHK>(G, D, 2, x, y - 1, z), hbar<CHK>(G, D, 2, x - 1, y - 1, z), 0, x);
  float lxy = bea(G, lx, lxm, 1, y);
  float hx = bea(G, hbar<CHK>(G, D, 2, x, y, z + 1), hbar<CHK>(G, D, 2, x - 1, y, z + 1), 0, x);
  float hxm = bea(G, hbar<CHK>(G, D, 2, x, y - 1, z + 1), hbar<CHK>(G, D, 2, x - 1, y - 1, z + 1), 0, x);
  float hxy = bea(G, hx, hxm, 1, y);
  Hs[2] = (lxy + hxy) / 2.0f;
}
// Interior cells with exact interpolation: the same samples and the same arithmetic as colocate_t, but
// every address is one base offset plus a constant stride (no halo rule, no per-sample index products).
__device__ __forceinline__ void colocate_interior(const GridDev& G, const DetDev& D, int x, int y, int z, float* Es, float* Hs) {
  const long long N = (long long)G.nx * G.ny * G.nz;
  const long long sy_g = G.nz, sx_g = (long long)G.ny * G.nz;
  const float* e = G.E + ((long long)x * G.ny + y) * G.nz + z;
  const float* h = G.H + ((long long)x * G.ny + y) * G.nz + z;
  const int hy_ = D.hi[1] - D.lo[1] + 1, hz_ = D.hi[2] - D.lo[2] + 1, hx_ = D.hi[0] - D.lo[0] + 1;
  const long long hn = (long long)hx_ * hy_ * hz_, sy_h = hz_, sx_h = (long long)hy_ * hz_;
  const float* p = D.hprev + ((long long)(x - D.lo[0] + 1) * hy_ + (y - D.lo[1] + 1)) * hz_ + (z - D.lo[2]);
#define HB(c, off_h, off_g) ((p[(c) * hn + (off_h)] + h[(c) * N + (off_g)]) / 2.0f)
  float lo = bea(G, e[0], e[-sx_g], 0, x);
  float hi = bea(G, e[1], e[1 - sx_g], 0, x);
  Es[0] = (lo + hi) / 2.0f;
  lo = bea(G, e[N], e[N - sy_g], 1, y);
  hi = bea(G, e[N + 1], e[N + 1 - sy_g], 1, y);
  Es[1] = (lo + hi) / 2.0f;
  Es[2] = e[2 * N];
  Hs[0] = bea(G, HB(0, 0, 0), HB(0, -sy_h, -sy_g), 1, y);
  Hs[1] = bea(G, HB(1, 0, 0), HB(1, -sx_h, -sx_g), 0, x);
  const float lx = bea(G, HB(2, 0, 0), HB(2, -sx_h, -sx_g), 0, x);
  const float lxm = bea(G, HB(2, -sy_h, -sy_g), HB(2, -sx_h - sy_h, -sx_g - sy_g), 0, x);
  const float lxy = bea(G, lx, lxm, 1, y);
  const float hx = bea(G, HB(2, 1, 1), HB(2, 1 - sx_h, 1 - sx_g), 0, x);
  const float hxm = bea(G, HB(2, 1 - sy_h, 1 - sy_g), HB(2, 1 - sx_h - sy_h, 1 - sx_g - sy_g), 0, x);
  const float hxy = bea(G, hx, hxm, 1, y);
  Hs[2] = (lxy + hxy) / 2.0f;
#undef HB
}
__device__ __forceinline__ void colocate(const GridDev& G, const DetDev& D, int x, int y, int z, float* Es, float* Hs) {
  // the stencil reads x-1..x, y-1..y, z..z+1: away from the faces no halo rule applies
  if (x >= 1 && x < G.nx && y >= 1 && y < G.ny && z >= 0 && z + 1 < G.nz) {
    if (D.flags & DET_EXACT) colocate_interior(G, D, x, y, z, Es, Hs);
    else colocate_t<false>(G, D, x, y, z, Es, Hs);
  } else {
    colocate_t<true>(G, D, x, y, z, Es, Hs);
  }
}

// One thread per region cell: sample, then write the per-type result (or stage it for a reduction).
__device__ __forceinline__ void det_sample_body(const GridDev& G, const DetDev& D, const int t, const long long cell);
template <bool FUSED_MEAN>
__device__ __forceinline__ void det_emit(const GridDev& G, const DetDev& D, const int t, const long long cell, const int rx, const int ry, const int rz,
                                         const int x, const int y, const int z, const float* Es, const float* Hs, float* e_out);
__global__ void det_sample_kernel(const GridDev G, const DetDev D, const int t) {
  const long long n = (long long)(D.hi[0] - D.lo[0]) * (D.hi[1] - D.lo[1]) * (D.hi[2] - D.lo[2]);
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= n) return;
  det_sample_body(G, D, t, cell);
}
__global__ void det_sample_batch_kernel(const GridDev G, const DetDev* __restrict__ dets, const int t, const int inverse) {
  __shared__ DetDev sD;
  det_stage_descriptor(&sD, dets + blockIdx.y);
  const DetDev& D = sD;
  if (((D.flags & DET_INVERSE) != 0) != (inverse != 0) || !D.on[t] || (D.flags & DET_VOLUME)) return;
  const long long n = (long long)(D.hi[0] - D.lo[0]) * (D.hi[1] - D.lo[1]) * (D.hi[2] - D.lo[2]);
  for (long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x; cell < n; cell += (long long)gridDim.x * blockDim.x)
    det_sample_body(G, D, t, cell);
}
__device__ __forceinline__ void det_sample_body(const GridDev& G, const DetDev& D, const int t, const long long cell) {
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const long long n = (long long)ex * ey * ez;
  int rx, ry, rz;
  if (n < 0x7fffffffLL) {
    const unsigned u = (unsigned)cell;
    rz = (int)(u % (unsigned)ez);
    ry = (int)((u / (unsigned)ez) % (unsigned)ey);
    rx = (int)(u / ((unsigned)ez * (unsigned)ey));
  } else {
    rz = (int)(cell % ez);
    ry = (int)((cell / ez) % ey);
    rx = (int)(cell / ((long long)ez * ey));
  }
  const int x = D.lo[0] + rx, y = D.lo[1] + ry, z = D.lo[2] + rz;
  // energy.py:118-143 with as_slices and no averaging keeps three planes of the region: O(surface) work
  if ((D.flags & DET_SLICES) && !(D.flags & DET_SLICE_MEAN) && rx != D.slice_idx[0] && ry != D.slice_idx[1] && rz != D.slice_idx[2]) return;
  if (D.flags & DET_CLOSED) {  // only the shell contributes (interior scratch cells stay at their initial zero)
    const bool fx = ((D.aux >> 0) & 1) && (rx == 0 || rx == ex - 1), fy = ((D.aux >> 1) & 1) && (ry == 0 || ry == ey - 1);
    const bool fz = ((D.aux >> 2) & 1) && (rz == 0 || rz == ez - 1);
    if (!fx && !fy && !fz) return;
  }
  float Es[3], Hs[3];
  colocate(G, D, x, y, z, Es, Hs);
  det_emit<false>(G, D, t, cell, rx, ry, rz, x, y, z, Es, Hs, nullptr);
}
// Per-type Detector.update on the co-located sample of one cell (objects/detectors/*.py): writes the
// state entry or stages the value for a reduction.  FUSED_MEAN: the caller reduces the slice means
// itself (det_volume.cuh) and only wants the cell's energy back through e_out.
template <bool FUSED_MEAN>
__device__ __forceinline__ void det_emit(const GridDev& G, const DetDev& D, const int t, const long long cell, const int rx, const int ry, const int rz,
                                         const int x, const int y, const int z, const float* Es, const float* Hs, float* e_out) {
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const long long n = (long long)ex * ey * ez;
  const int slot = D.arr_idx[t];
  const bool staged = (D.flags & DET_REDUCE) || ((D.flags & DET_SLICES) && (D.flags & DET_SLICE_MEAN));
  if (D.kind == 3 && !staged) {
    // phasor accumulation (phasor.py:186-235): state += v * exp(i w t) * scale * window.  All read-modify-
    // write operands of the cell are loaded before the first store (the state may alias nothing the
    // compiler can prove, so interleaved loads / stores would serialise into one DRAM round trip each)
    float vals[6];
    int nci = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c)
      if (D.comp_mask & (1 << c)) vals[nci++] = (c < 3) ? Es[c] : Hs[c - 3];
    float2* __restrict__ st = reinterpret_cast<float2*>(D.state[0]);
    const float w = D.window[t];
    for (int f = 0; f < D.nf; ++f) {
      const float2 ph = D.ph_table[(long long)t * D.nf + f];
      float2 acc[6];
#pragma unroll
      for (int ci = 0; ci < 6; ++ci)
        if (ci < nci) acc[ci] = st[((long long)f * D.ncomp + ci) * n + cell];
#pragma unroll
      for (int ci = 0; ci < 6; ++ci) {
        if (ci >= nci) continue;
        const float re = ((vals[ci] * ph.x) * D.scale) * w;
        const float im = ((vals[ci] * ph.y) * D.scale) * w;
        if (D.flags & DET_INVERSE) { acc[ci].x -= re; acc[ci].y -= im; } else { acc[ci].x += re; acc[ci].y += im; }
      }
#pragma unroll
      for (int ci = 0; ci < 6; ++ci)
        if (ci < nci) st[((long long)f * D.ncomp + ci) * n + cell] = acc[ci];
    }
  } else if (D.kind == 0 || D.kind == 3) {  // field / phasor: selected components
    int ci = 0;
    for (int c = 0; c < 6; ++c) {
      if (!(D.comp_mask & (1 << c))) continue;
      const float v = (c < 3) ? Es[c] : Hs[c - 3];
      if (staged) {
        D.scratch[ci * n + cell] = v;
      } else if (D.kind == 0) {
        D.state[0][((long long)slot * D.ncomp + ci) * n + cell] = v;
      } else {
        float2* st = reinterpret_cast<float2*>(D.state[0]);
        const float w = D.window[t];
        for (int f = 0; f < D.nf; ++f) {
          const float2 ph = D.ph_table[(long long)t * D.nf + f];
          const float re = ((v * ph.x) * D.scale) * w;
          const float im = ((v * ph.y) * D.scale) * w;
          float2 s = st[((long long)f * D.ncomp + ci) * n + cell];
          if (D.flags & DET_INVERSE) { s.x -= re; s.y -= im; } else { s.x += re; s.y += im; }
          st[((long long)f * D.ncomp + ci) * n + cell] = s;
        }
      }
      ++ci;
    }
  } else if (D.kind == 1) {  // energy (metrics.py:55-67)
    const long long N = (long long)G.nx * G.ny * G.nz;
    const long long gidx = ((long long)x * G.ny + y) * G.nz + z;
    float eE = 0.0f, eH = 0.0f;
    if (G.eps_tier == 9 || G.mu_tier == 9) {
      // full-tensor media (metrics.py:37-53): eps = inv(inv_eps), energy = 0.5 E.eps.E + 0.5 H.mu.H
      float A[2][3][3], Minv[2][3][3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          A[0][i][j] = (G.eps_tier == 9) ? G.eps[(long long)(3 * i + j) * G.eps_cs + gidx] : (i == j ? G.eps[(long long)(G.eps_tier == 1 ? 0 : i) * G.eps_cs + gidx] : 0.0f);
          A[1][i][j] = (G.mu_tier == 9) ? G.mu[(long long)(3 * i + j) * G.mu_cs + gidx]
                                        : (i == j ? (G.mu ? G.mu[(long long)(G.mu_tier == 1 ? 0 : i) * G.mu_cs + gidx] : G.inv_mu_scalar) : 0.0f);
        }
      for (int q = 0; q < 2; ++q) {
        const float(*M)[3] = A[q];
        const float c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1], c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2], c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
        const float det = (M[0][0] * c00 + M[0][1] * c01) + M[0][2] * c02;
        const float id = 1.0f / det;
        Minv[q][0][0] = c00 * id; Minv[q][1][0] = c01 * id; Minv[q][2][0] = c02 * id;
        Minv[q][0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id;
        Minv[q][1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id;
        Minv[q][2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id;
        Minv[q][0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
        Minv[q][1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
        Minv[q][2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
      }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          eE += (Es[i] * Minv[0][i][j]) * Es[j];
          eH += (Hs[i] * Minv[1][i][j]) * Hs[j];
        }
      eE = 0.5f * eE;
      eH = 0.5f * eH;
    } else {
      for (int c = 0; c < 3; ++c) {
        const float ie = G.eps[c * G.eps_cs + gidx];
        const float im = G.mu ? G.mu[c * G.mu_cs + gidx] : G.inv_mu_scalar;
        const float a = (0.5f * (1.0f / ie)) * (fabsf(Es[c]) * fabsf(Es[c]));
        const float b = (0.5f * (1.0f / im)) * (fabsf(Hs[c]) * fabsf(Hs[c]));
        eE = (c == 0) ? a : eE + a;
        eH = (c == 0) ? b : eH + b;
      }
    }
    (void)N;
    const float e = eE + eH;
    if (FUSED_MEAN) {
      *e_out = e;
    } else if (staged) {
      D.scratch[cell] = e;
    } else if (D.flags & DET_SLICES) {
      if (rz == D.slice_idx[2]) D.state[0][((long long)slot * ex + rx) * ey + ry] = e;
      if (ry == D.slice_idx[1]) D.state[1][((long long)slot * ex + rx) * ez + rz] = e;
      if (rx == D.slice_idx[0]) D.state[2][((long long)slot * ey + ry) * ez + rz] = e;
    } else {
      D.state[0][(long long)slot * n + cell] = e;
    }
  } else {  // Poynting (metrics.py:99-117, poynting_flux.py:171-195)
    float S[3] = {Es[1] * Hs[2] - Es[2] * Hs[1], Es[2] * Hs[0] - Es[0] * Hs[2], Es[0] * Hs[1] - Es[1] * Hs[0]};
    if (D.flags & DET_NEGATIVE) { S[0] = -S[0]; S[1] = -S[1]; S[2] = -S[2]; }
    if (D.flags & DET_CLOSED) {
      // net_poynting_flux_through_box (metrics.py:120-160): +S_a * area on the max face, -S_a * area on
      // the min face of every active axis; the per-cell terms are summed by det_reduce_all_kernel
      const int r[3] = {rx, ry, rz}, e3[3] = {ex, ey, ez};
      float c = 0.0f;
      for (int a = 0; a < 3; ++a) {
        if (!((D.aux >> a) & 1)) continue;
        const float sw = S[a] * D.weights[a * n + cell];
        if (r[a] == e3[a] - 1) c = c + sw;
        if (r[a] == 0) c = c - sw;
      }
      D.scratch[cell] = c;
    } else if (D.flags & DET_KEEP_ALL) {
      for (int c = 0; c < 3; ++c) {
        if (staged) D.scratch[c * n + cell] = S[c];
        else D.state[0][((long long)slot * 3 + c) * n + cell] = S[c];
      }
    } else {
      if (staged) D.scratch[cell] = S[D.aux];
      else D.state[0][(long long)slot * n + cell] = S[D.aux];
    }
  }
}

// Deterministic block reduction: value v of block v = sum over cells of scratch[v][cell] * w[cell].
__global__ void det_reduce_all_kernel(const DetDev D, const int t, const int nvals, const int w_per_val) {
  const int v = blockIdx.x;
  const long long n = (long long)(D.hi[0] - D.lo[0]) * (D.hi[1] - D.lo[1]) * (D.hi[2] - D.lo[2]);
  __shared__ float sm[1024];
  float acc = 0.0f;
  const float* w = (D.weights && !(D.flags & DET_CLOSED)) ? D.weights + (w_per_val ? (long long)v * n : 0) : nullptr;
  for (long long c = threadIdx.x; c < n; c += blockDim.x) {
    const float val = D.scratch[v * n + c];
    acc += w ? val * w[c] : val;
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const float r = sm[0];
  const int slot = D.arr_idx[t];
  if (D.kind == 0) {
    D.state[0][(long long)slot * D.ncomp + v] = r / D.wsum;
  } else if (D.kind == 1) {
    D.state[0][slot] = r;
  } else if (D.kind == 2) {
    D.state[0][(long long)slot * nvals + v] = r;
  } else {
    const float mean = r / D.wsum;
    float2* st = reinterpret_cast<float2*>(D.state[0]);
    const float wt = D.window[t];
    for (int f = 0; f < D.nf; ++f) {
      const float2 ph = D.ph_table[(long long)t * D.nf + f];
      const float re = ((mean * ph.x) * D.scale) * wt;
      const float im = ((mean * ph.y) * D.scale) * wt;
      float2 s = st[(long long)f * D.ncomp + v];
      if (D.flags & DET_INVERSE) { s.x -= re; s.y -= im; } else { s.x += re; s.y += im; }
      st[(long long)f * D.ncomp + v] = s;
    }
  }
}

// Energy slice means (energy.py:112-116): out = mean of the staged energy over one axis.
__global__ void det_slice_mean_kernel(const DetDev D, const int t, const int axis) {
  const int ex = D.hi[0] - D.lo[0], ey = D.hi[1] - D.lo[1], ez = D.hi[2] - D.lo[2];
  const int dims[3] = {ex, ey, ez};
  const int a1 = (axis == 0) ? 1 : 0, a2 = (axis == 2) ? 1 : 2;  // kept axes, ascending
  const long long nout = (long long)dims[a1] * dims[a2];
  const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (o >= nout) return;
  int idx[3];
  idx[a1] = (int)(o / dims[a2]);
  idx[a2] = (int)(o % dims[a2]);
  float acc = 0.0f;
  for (int q = 0; q < dims[axis]; ++q) {
    idx[axis] = q;
    acc += D.scratch[((long long)idx[0] * ey + idx[1]) * ez + idx[2]];
  }
  const int slot = D.arr_idx[t];
  const int key = (axis == 2) ? 0 : (axis == 1 ? 1 : 2);  // XY <- mean over z, XZ <- y, YZ <- x
  D.state[key][(long long)slot * nout + o] = acc / (float)dims[axis];
}

// ------------------------------------------------------------------------------------------------
// low-precision conversions for the recorder (DtypeConversion, modules.py:138-161)
// ------------------------------------------------------------------------------------------------
template <int EB, int MB, int BIAS, int MODE>  // MODE 0: e4m3fnuz, 1: e4m3fn, 2: e5m2 (IEEE-like)
__device__ __forceinline__ uint8_t f32_to_f8(float f) {
  const uint32_t sign = (__float_as_uint(f) >> 31) & 1u;
  const uint8_t nan_code = (MODE == 0) ? 0x80 : (uint8_t)((sign << 7) | (MODE == 1 ? 0x7f : 0x7e));
  if (isnan(f)) return nan_code;
  const float a = fabsf(f);
  const float max_val = (MODE == 0) ? 240.0f : (MODE == 1 ? 448.0f : 57344.0f);
  if (isinf(a)) return (MODE == 2) ? (uint8_t)((sign << 7) | 0x7c) : nan_code;
  int ex;
  (void)frexpf(a, &ex);  // a = m * 2^ex, m in [0.5, 1)
  int e_unb = ex - 1;    // floor(log2 a)
  const int e_min = 1 - BIAS;
  if (a == 0.0f) return (MODE == 0) ? 0 : (uint8_t)(sign << 7);
  if (e_unb < e_min) e_unb = e_min;
  const float quantum = ldexpf(1.0f, e_unb - MB);
  float r = rintf(a / quantum) * quantum;  // round to nearest even on the target grid
  if (r > max_val) return (MODE == 2) ? (uint8_t)((sign << 7) | 0x7c) : nan_code;
  if (r == 0.0f) return (MODE == 0) ? 0 : (uint8_t)(sign << 7);
  (void)frexpf(r, &ex);
  e_unb = ex - 1;
  uint32_t ebits, mbits;
  if (e_unb < e_min) {
    ebits = 0;
    mbits = (uint32_t)(r / ldexpf(1.0f, e_min - MB));
  } else {
    ebits = (uint32_t)(e_unb + BIAS);
    mbits = (uint32_t)((r / ldexpf(1.0f, e_unb) - 1.0f) * (float)(1 << MB));
  }
  return (uint8_t)((sign << 7) | (ebits << MB) | mbits);
}

template <int EB, int MB, int BIAS, int MODE>
__device__ __forceinline__ float f8_to_f32(uint8_t b) {
  const uint32_t sign = b >> 7;
  const uint32_t ebits = (b >> MB) & ((1u << EB) - 1u);
  const uint32_t mbits = b & ((1u << MB) - 1u);
  if (MODE == 0 && b == 0x80) return __uint_as_float(0x7fc00000u);
  if (MODE == 1 && ebits == 15 && mbits == 7) return __uint_as_float(0x7fc00000u);
  if (MODE == 2 && ebits == 31) return mbits ? __uint_as_float(0x7fc00000u) : (sign ? -INFINITY : INFINITY);
  float v = (ebits == 0) ? ldexpf((float)mbits, 1 - BIAS - MB) : ldexpf(1.0f + (float)mbits / (float)(1 << MB), (int)ebits - BIAS);
  return sign ? -v : v;
}

__device__ __forceinline__ void rec_store(void* data, int dtype, long long idx, float v) {
  switch (dtype) {
    case 0: reinterpret_cast<float*>(data)[idx] = v; break;
    case 1: reinterpret_cast<__nv_bfloat16*>(data)[idx] = __float2bfloat16_rn(v); break;
    case 2: reinterpret_cast<__half*>(data)[idx] = __float2half_rn(v); break;
    case 3: reinterpret_cast<uint8_t*>(data)[idx] = f32_to_f8<4, 3, 8, 0>(v); break;
    case 4: reinterpret_cast<uint8_t*>(data)[idx] = f32_to_f8<4, 3, 7, 1>(v); break;
    default: reinterpret_cast<uint8_t*>(data)[idx] = f32_to_f8<5, 2, 15, 2>(v); break;
  }
}
__device__ __forceinline__ float rec_load(const void* data, int dtype, long long idx) {
  switch (dtype) {
    case 0: return reinterpret_cast<const float*>(data)[idx];
    case 1: return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(data)[idx]);
    case 2: return __half2float(reinterpret_cast<const __half*>(data)[idx]);
    case 3: return f8_to_f32<4, 3, 8, 0>(reinterpret_cast<const uint8_t*>(data)[idx]);
    case 4: return f8_to_f32<4, 3, 7, 1>(reinterpret_cast<const uint8_t*>(data)[idx]);
    default: return f8_to_f32<5, 2, 15, 2>(reinterpret_cast<const uint8_t*>(data)[idx]);
  }
}

struct RecPlane {
  int lo[3], hi[3];  // 1-cell-thick interface plane (boundary.py:117-144), local coordinates
  void* data[2];     // [E, H] recorder buffers (slots, 3, *face)
};
#define FDTDX_MAX_PML 6
struct RecDev {
  int n_planes;
  int dtype;
  RecPlane planes[FDTDX_MAX_PML];
};

// blockIdx.y = plane * 2 + field
// a face holds < 2^31 / 3 cells: 32-bit index arithmetic (64-bit div / mod is ~100 instructions each)
__global__ void rec_record_kernel(const RecDev R, float* E, float* H, int nx, int ny, int nz, int slot) {
  pdl_trigger();
  const RecPlane& pl = R.planes[blockIdx.y >> 1];
  const int fld = blockIdx.y & 1;
  const float* F = fld ? H : E;
  const int ex = pl.hi[0] - pl.lo[0], ey = pl.hi[1] - pl.lo[1], ez = pl.hi[2] - pl.lo[2];
  const unsigned fn = (unsigned)ex * ey * ez;
  const long long N = (long long)nx * ny * nz;
  pdl_wait();
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 3 * fn; idx += gridDim.x * blockDim.x) {
    const int c = (int)(idx / fn);
    unsigned r = idx - c * fn;
    const int z = pl.lo[2] + (int)(r % (unsigned)ez); r /= (unsigned)ez;
    const int y = pl.lo[1] + (int)(r % (unsigned)ey);
    const int x = pl.lo[0] + (int)(r / (unsigned)ey);
    const float v = F[c * N + ((long long)x * ny + y) * nz + z];
    rec_store(pl.data[fld], R.dtype, (long long)slot * 3 * fn + idx, v);
  }
}

__global__ void rec_replay_kernel(const RecDev R, float* E, float* H, int nx, int ny, int nz, int sa, int sb, float w) {
  pdl_trigger();
  const RecPlane& pl = R.planes[blockIdx.y >> 1];
  const int fld = blockIdx.y & 1;
  float* F = fld ? H : E;
  const int ex = pl.hi[0] - pl.lo[0], ey = pl.hi[1] - pl.lo[1], ez = pl.hi[2] - pl.lo[2];
  const unsigned fn = (unsigned)ex * ey * ez;
  const long long N = (long long)nx * ny * nz;
  pdl_wait();
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 3 * fn; idx += gridDim.x * blockDim.x) {
    const int c = (int)(idx / fn);
    unsigned r = idx - c * fn;
    const int z = pl.lo[2] + (int)(r % (unsigned)ez); r /= (unsigned)ez;
    const int y = pl.lo[1] + (int)(r % (unsigned)ey);
    const int x = pl.lo[0] + (int)(r / (unsigned)ey);
    float v = rec_load(pl.data[fld], R.dtype, (long long)sa * 3 * fn + idx);
    if (sb != sa) {
      const float nxt = rec_load(pl.data[fld], R.dtype, (long long)sb * 3 * fn + idx);
      v = v + w * (nxt - v);  // time_filter.py:237-250
    }
    F[c * N + ((long long)x * ny + y) * nz + z] = v;
  }
}

struct BoxList {
  int n;
  int lo[FDTDX_MAX_PML][3], hi[FDTDX_MAX_PML][3];
};
// zero E and H inside every PML slab (backward.py:117-122)
__global__ void reset_pml_kernel(const BoxList B, float* E, float* H, int nx, int ny, int nz) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int ex = B.hi[b][0] - B.lo[b][0], ey = B.hi[b][1] - B.lo[b][1], ez = B.hi[b][2] - B.lo[b][2];
  const long long n = (long long)ex * ey * ez;
  const long long N = (long long)nx * ny * nz;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int z = B.lo[b][2] + (int)(r % ez); r /= ez;
    const int y = B.lo[b][1] + (int)(r % ey);
    const int x = B.lo[b][0] + (int)(r / ey);
    const long long g = ((long long)x * ny + y) * nz + z;
    for (int c = 0; c < 3; ++c) {
      E[c * N + g] = 0.0f;
      H[c * N + g] = 0.0f;
    }
  }
}

// Pack / unpack the two tangential components of one x plane (x-slab halo).
__global__ void halo_pack_kernel(const float* F, float* out, int nx, int ny, int nz, int plane_x) {
  const long long pn = (long long)ny * nz;
  const long long N = pn * nx;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < 2 * pn; idx += (long long)gridDim.x * blockDim.x) {
    const int c = 1 + (int)(idx / pn);
    out[idx] = F[c * N + (long long)plane_x * pn + (idx % pn)];
  }
}

// ------------------------------------------------------------------------------------------------
// Total field energy (core/physics/metrics.py:55-67, diagonal tiers), the per-step global reduction
// of EnergyThresholdCondition (fdtd/stop_conditions.py:81-147).  Per cell, in float32 and in the
// reference's op order: 0.5 * (1 / inv_eps_c) * E_c^2 summed over c as (x + y) + z, likewise for H,
// then eE + eH.  The sum over cells is carried in float64 with a fixed launch geometry, so the result
// is deterministic; stage 2 (one block) folds the per-block partials and writes one float.
// ------------------------------------------------------------------------------------------------
#define FDTDX_ENERGY_BLOCKS (148 * 4)
// zpad: the last zpad cells of every z row are padding of the caller's grid (zero fields, zero inv_eps): not part of the sum
__global__ void __launch_bounds__(256) energy_partial_kernel(const GridDev G, double* __restrict__ partial, const int zpad) {
  const long long N = (long long)G.nx * G.ny * G.nz;
  double acc = 0.0;
  for (long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x; cell < N; cell += (long long)gridDim.x * blockDim.x) {
    if (zpad > 0 && (int)(cell % G.nz) >= G.nz - zpad) continue;
    float eE[3], eH[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e = G.E[c * N + cell], h = G.H[c * N + cell];
      const float ie = G.eps[c * G.eps_cs + cell];
      const float im = G.mu ? G.mu[c * G.mu_cs + cell] : G.inv_mu_scalar;
      eE[c] = (0.5f * (1.0f / ie)) * (e * e);
      eH[c] = (0.5f * (1.0f / im)) * (h * h);
    }
    const float cellE = ((eE[0] + eE[1]) + eE[2]) + ((eH[0] + eH[1]) + eH[2]);
    acc += (double)cellE;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) energy_final_kernel(const double* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)sh[0];
}
