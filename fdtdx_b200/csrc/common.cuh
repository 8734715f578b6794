// Shared device-side descriptors for the Yee hot-path kernels (sm_100a).
//
// Arithmetic discipline: this translation unit is compiled with -fmad=false and IEEE division so
// that every float32 expression rounds exactly like the reference's jnp expression evaluated
// op-by-op; the kernels are HBM-bound, so un-fused FMUL+FADD is free.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define FDTDX_MAX_SRC 8
#define FDTDX_MAX_WALL 12
#define FDTDX_MAX_HPBOX 4

struct AxisPmlDev {
  int lo_len;    // local cells [0, lo_len) belong to the '-' slab (0: none)
  int hi_start;  // local cells [hi_start, n) belong to the '+' slab (n: none)
  int hi_len;
  int kappa_one; // both slabs have kappa == 1 (correction == psi, perfectly_matched_layer.py:184)
  int vec_ok;    // z axis only: slab rows are 16-byte aligned (thickness % 4 == 0, psi pointers aligned)
  // per-cell coefficient tables along this axis, local length n; zero outside the slabs.
  // k* = 1/kappa - 1.  E-side tables feed curl_H (E update), H-side feed curl_E (H update).
  const float *aE, *bE, *kE, *aH, *bH, *kH;
  float* psiE[2][2];  // [side: 0 lo, 1 hi][which: psi_1, psi_2]
  float* psiH[2][2];
};

struct WallDev {
  int kind;  // 0 PEC (E), 1 PMC (H)
  int axis;
  int lo[3], hi[3];
};

struct SrcDev {
  int kind;  // 0 TFSF plane, 1 dipole
  int lo[3], hi[3];
  int normal_axis;
  float sign;  // +1 / -1 (direction); negated for the reverse pass at launch
  int profile_kind;
  float p[6];  // profile params, float32-rounded exactly like the weak python scalars
  float static_amp;
  float cE, cH;
  const float *Einc, *Hinc, *toffE, *toffH;  // (3, face)
  // complex (lossy-mode) profiles: imaginary parts or nullptr, injected in quadrature with the carrier
  // phase shifted by -pi/2 (tfsf.py:266-283, 366-383); pq = float32(phase_shift - pi/2)
  const float *EincI, *HincI;
  float pq;
  const float* signal;
  int signal_len;
  const float* hfilter;
  int hfilter_len;
  const uint8_t* on;   // per-time-step gate or nullptr (default always-on switch)
  const float* t_adj;  // remapped time step (source.py:44-49) or nullptr
  // dipole
  int pol, electric;
  float dip_scale;
};

struct StepParams {
  int nx, ny, nz;
  int wrap[3];
  int sym[3];  // config.symmetry axis: the MIN-side halo is never wrapped (update.py:121-125); the far side still is
  int x_lo_mode, x_hi_mode;  // 0 zero halo, 1 local wrap, 2 neighbour halo buffer
  float cour, eta0, inv_mu_scalar, dt;
  float* E;
  float* H;
  // optional separate destinations of the E / H half-step (staged kernels only; nullptr = in place):
  // the adjoint pass recomputes a step into scratch without copying the state first
  float* E_out;
  float* H_out;
  const float* eps;
  const float* mu;  // nullptr => scalar
  const float* sigE;
  const float* sigH;
  long long eps_cs, mu_cs, sigE_cs, sigH_cs;  // component strides (0 for 1-component tiers)
  const float* sB[3];
  const float* sF[3];
  AxisPmlDev pml[3];
  int simulate;
  int psi_store;  // 0: compute psi' for the correction but leave the stored psi untouched (adjoint recompute)
  int p_store;    // 0: likewise leave the ADE polarisation buffers untouched
  int n_walls;
  const WallDev* walls;  // device array
  int n_src;
  const SrcDev* src;  // device array
  // hot-loop copies in kernel-parameter (constant) space: wall boxes, source bounding boxes and the
  // union x-range of all sources (a CTA-uniform test that keeps the injection path cold)
  WallDev wallp[FDTDX_MAX_WALL];
  int src_lo[FDTDX_MAX_SRC][3], src_hi[FDTDX_MAX_SRC][3];
  int src_x0, src_x1;
  int src_inline;  // 1: the half-step kernels inject in their cold pass; 0: src_apply_kernel does (broad sources)
  int wall_x0[2], wall_x1[2];  // union x-range of the PEC [0] / PMC [1] walls
  // ADE (update.py:316-350)
  int n_poles, has_c4;
  const float* P_cur;  // dispersive_P_curr
  float* P_new;        // buffer that held dispersive_P_prev; receives the new P_curr
  const float *c1, *c2, *c3, *c4;
  long long c_cs;  // coefficient component stride (0: isotropic, N: per-axis)
  // x-slab neighbours (x_lo_mode / x_hi_mode == 2): Hy,Hz of plane x0-1 and Ey,Ez of plane x1.  Either a
  // packed (2,ny,nz) staging buffer filled by an exchange (component stride = ny*nz), or the
  // neighbour rank's own field array mapped over NVLink (component stride = that rank's Nx*ny*nz).
  const float* haloH;
  const float* haloE;
  long long haloH_cs, haloE_cs;
  int xchunk;
  int flat_lz;         // staged kernels with flat tiles (yee_tma.cuh, TmaRt): z quads per row (Nz / 4); 0 otherwise
  int x_begin, x_end;  // plane range of this launch (sub-ranges let the halo exchange overlap)
  // Peer-memory halo (x-slab sharding, neighbour arrays mapped over NVLink): producer / consumer order
  // with the neighbour rank is kept INSIDE the half-step kernel.  Only the CTAs of the boundary chunk
  // (E step: the chunk holding plane 0; H step: the chunk holding plane nx-1) poll the neighbour's
  // progress counter before they touch the shared plane, and the last of their warps to finish
  // publishes this rank's counter; every other CTA runs unordered, so the interior overlaps the wait.
  const int* peer_wait;   // neighbour's counter (nullptr: nothing to wait for)
  int peer_wait_target;
  int* peer_signal;       // this rank's counter (nullptr: nobody listens)
  int peer_signal_value;
  int* peer_ctr;          // arrivals of the boundary chunk's warps (reset by the last one)
  int peer_total;
  int* peer_err;          // set to 1 when a wait gave up (neighbour stalled): results are invalid
  int z_reverse;          // E step: blockIdx.z counts chunks from the high-x end, so the boundary chunk runs last
  // Bloch-periodic axes with a non-zero wave vector (bloch.py:61-96, complex fields): the run is two
  // real systems (Re, Im) stepped side by side.  A wrapped ghost value on the LOW side of axis a is
  // F_ghost = F[N-1] * conj(phase): self * bc[a] + partner * bs[a]; on the HIGH side F[0] * phase:
  // self * bc[a] - partner * bs[a], where (bc, bs) = (cos, sin) of k_a * L_a for the Re system and
  // (cos, -sin) for the Im system, and bH / bE are the partner system's H / E arrays (nullptr: plain wrap).
  const float* bH;
  const float* bE;
  float bc[3], bs[3];
  // Forward H half-step on a step where a large exact-interpolation detector is on: the kernel also
  // stores the H it loaded (H before this update = H_prev of update.py:1088) into this (3,Nx,Ny,Nz)
  // scratch, so the detector pass needs no separate H_prev copy (det_volume.cuh).  nullptr: off.
  float* hprev_out;
  // planes / rows the active detectors read (incl. their x-1 / y-1 halo): {x0, x1, y0, y1} per box
  int hprev_nbox;
  int hprev_box[FDTDX_MAX_HPBOX][4];
};
__device__ __forceinline__ bool hprev_wanted(const StepParams& P, const int i, const int j) {
  bool w = false;
#pragma unroll
  for (int q = 0; q < FDTDX_MAX_HPBOX; ++q)
    if (q < P.hprev_nbox && i >= P.hprev_box[q][0] && i < P.hprev_box[q][1] && j >= P.hprev_box[q][2] && j < P.hprev_box[q][3]) w = true;
  return w;
}

// Programmatic dependent launch (every per-step kernel is launched with the programmatic-stream-
// serialization attribute): pdl_trigger() lets the next kernel of the stream begin its prologue as soon as
// all CTAs of this grid have started; pdl_wait() returns once the previous kernel of the stream has
// completed and its writes are visible.  Everything before pdl_wait() may touch only constant tables.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int V>
struct Vec {
  float v[V];
};

// FDTDX_RAGGED (per translation unit): the 4-cells-per-thread marching kernels on grids whose rows
// are not 16-byte aligned (Nz % 4 != 0 or unaligned buffers).  A warp still owns 128 consecutive z
// cells of a row, but INTERLEAVED: lane l holds cells l, l+32, l+64, l+96 of the tile (element stride
// FDTDX_ES = 32), so every 32-bit access of the warp is one fully coalesced 128-byte line and all
// per-thread bookkeeping is still amortised over four cells.  nv = valid elements of this thread;
// cells beyond the row read as 0 and are never written.
#ifndef FDTDX_RAGGED
#define FDTDX_RAGGED 0
#endif
#define FDTDX_ES (FDTDX_RAGGED ? 32 : 1)

template <int V>
__device__ __forceinline__ Vec<V> ldv(const float* __restrict__ p, const int nv = V) {
  Vec<V> r;
  if constexpr (V == 4 && !FDTDX_RAGGED) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < V; ++e) r.v[e] = (e < nv) ? p[e * FDTDX_ES] : 0.0f;
  }
  return r;
}
template <int V>
__device__ __forceinline__ void stv(float* __restrict__ p, const Vec<V>& r, const int nv = V) {
  if constexpr (V == 4 && !FDTDX_RAGGED) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
#pragma unroll
    for (int e = 0; e < V; ++e)
      if (e < nv) p[e * FDTDX_ES] = r.v[e];
  }
}
// Bloch ghost value: self * c + partner * s (see StepParams::bH)
template <int V>
__device__ __forceinline__ Vec<V> bloch_mix(const Vec<V>& a, const Vec<V>& b, const float c, const float s) {
  Vec<V> r;
#pragma unroll
  for (int e = 0; e < V; ++e) r.v[e] = a.v[e] * c + b.v[e] * s;
  return r;
}
template <int V>
__device__ __forceinline__ Vec<V> zerov() {
  Vec<V> r;
#pragma unroll
  for (int e = 0; e < V; ++e) r.v[e] = 0.f;
  return r;
}
