// C ABI implementation (include/fdtdx_b200.h): plan = constant tables + scratch; caller owns buffers.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda.h>
#include <cstdlib>
#include <array>
#include <map>
#include <tuple>

#include "../../include/fdtdx_b200.h"
#include "tma_cfg.h"
#include "aux_kernels.cuh"
#include "det_volume.cuh"
#include "common.cuh"
#include "tensor_kernels.cuh"
#include "adjoint_kernels.cuh"
#include "source_kernels.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(x)                                                                        \
  do {                                                                                     \
    cudaError_t _e = (x);                                                                  \
    if (_e != cudaSuccess)                                                                 \
      return fail(FDTDX_ECUDA, std::string(#x) + ": " + cudaGetErrorString(_e));           \
  } while (0)

static const double kEta0 = (4e-7 * M_PI) * 299792458.0;
#define MAX_IDX 512

struct PmlHost {
  int axis, dir, lo, hi, kappa_one;
  int lo_true, hi_true;  // the reference's slab (recorder interface plane, field reset); [lo, hi) may be a padded superset
  std::vector<float> aE, bE, kE, aH, bH, kH;
};
struct SrcHost {
  SrcDev d;
  std::vector<uint8_t> on;
};
struct DetHost {
  DetDev d;
  std::vector<uint8_t> on;
  int nvals;
  bool volume = false;  // DET_VOLUME requested and structurally possible (row-marching kernels, det_volume.cuh)
};

struct FdtdxPlan {
  int nx, ny, nz, xoff, nxg, T;
  double courant, dt;
  int eps_tier, mu_tier, sigE_tier, sigH_tier;
  double inv_mu_scalar;
  int wrap[3];
  bool metric;
  float* d_sB[3];
  float* d_sF[3];
  float* d_w[3];
  std::vector<PmlHost> pmls;
  std::vector<WallDev> walls;
  std::vector<SrcHost> srcs;
  SrcDev* d_srcs = nullptr;
  WallDev* d_walls = nullptr;
  std::vector<DetHost> dets;
  std::vector<DetDev> h_dets;
  DetDev* d_dets = nullptr;   // device copy of the detector descriptors (batched launches)
  bool dets_dirty = true;
  long long det_max_cells = 0, det_max_halo = 0;
  // recorder
  bool has_rec = false;
  int rec_dtype = 0, rec_slots = 0;
  std::vector<int32_t> slot_of_time, replay_a, replay_b;
  std::vector<float> replay_w;
  // dispersion
  int n_poles = 0, coeff_tier = 1, has_c4 = 0;
  int halo_lo = 0, halo_hi = 0;
  bool halo_lo_planned = false;  // x_offset > 0: a lower neighbour rank exists (known at plan_create, before halo_bind)
  void* slots[FDTDX_SLOT_COUNT][MAX_IDX];
  std::vector<void*> owned;
  bool finalized = false;
  AxisPmlDev axis[3];
  int p_parity = 0, e_parity = 0, h_parity = 0;
  long long launches = 0;
  int xchunk = 0, rows = 4;
  // peer-memory halo (NVLink): neighbour field arrays + progress flags mapped through CUDA IPC
  struct PeerLink {
    float* field = nullptr;  // lo: neighbour's H (3,nx_peer,ny,nz); hi: neighbour's E
    int* flags = nullptr;    // neighbour's {doneE, doneH} counters
    int nx = 0;
  } peer[2];
  int* d_flags = nullptr;          // this rank's {doneE, doneH}
  unsigned seqE = 0, seqH = 0;     // half-steps issued so far (identical on every rank)
  bool peer_mode = false;
  std::vector<std::pair<std::string, void*>> ipc_open;  // handle bytes -> mapped base
  int use_tma = -1;     // -1: decide from the environment (FDTDX_B200_TMA=0 disables), 0 / 1: forced
  int xchunk_tma = 0;   // planes per CTA of the TMA-staged kernels (0: heuristic)
  std::map<std::tuple<const void*, int, int>, CUtensorMap> tmaps;  // (base, components, box kind) -> map
  float* d_K = nullptr;  // tensor path: curl scratch (3,N)
  float *d_Etmp = nullptr, *d_Htmp = nullptr, *d_lamHx = nullptr, *d_ld = nullptr;  // adjoint scratch
  // fused adjoint kernel: second set of psi-cotangent buffers per half-step kind ([0] H, [1] E; index 2*q+w as
  // COT_PSI_*) and which set currently holds the value (1: the plan-owned one; copied back at the end of run_adjoint)
  std::vector<float*> d_cotpsi_alt[2];
  int cotpsi_parity[2] = {0, 0};
  // the detectors' cotangent kernels of one step are independent (atomic scatters): they run side by side on two
  // plan-owned streams forked from / joined to the caller's stream
  cudaStream_t adj_side[2] = {nullptr, nullptr};
  cudaEvent_t adj_fork = nullptr, adj_join[2] = {nullptr, nullptr};
  double* d_energy_partial = nullptr;  // total_energy: per-block partial sums
  bool adjoint_exact = false;          // run_adjoint_exact: VJP at the bound state, no reverse step
  // row-marching detector kernels (det_volume.cuh): plan-wide H_prev scratch in the layout of H, the
  // step whose forward H half-step fills it itself (StepParams::hprev_out), x planes per CTA
  float* d_hprev_full = nullptr;
  int hprev_fused_t = -1;
  int hprev_nbox = 0;
  int hprev_box[FDTDX_MAX_HPBOX][4];  // {x0, x1, y0, y1}: rows the detectors active at that step read
  int detv_xcl = 4;
  // Bloch axes with k != 0: this plan is one of the two real systems of a complex run
  int sym[3] = {0, 0, 0}, mirror[3] = {0, 0, 0};  // config.symmetry (see GridDev::sym)
  int zpad = 0;  // the last zpad z cells are padding of the caller's grid (plan.py): recorder planes keep the true Nz
  bool bloch = false;
  float bloch_c[3] = {1.f, 1.f, 1.f}, bloch_s[3] = {0.f, 0.f, 0.f};
};

extern "C" const char* fdtdx_b200_last_error(void) { return g_err.c_str(); }
extern "C" int fdtdx_b200_version(void) { return 100; }

template <typename T>
static int to_device(FdtdxPlan* p, const T* host, size_t n, T** out) {
  *out = nullptr;
  if (n == 0) return FDTDX_OK;
  CUDA_TRY(cudaMalloc((void**)out, n * sizeof(T)));
  p->owned.push_back(*out);
  if (host) CUDA_TRY(cudaMemcpy(*out, host, n * sizeof(T), cudaMemcpyHostToDevice));
  else CUDA_TRY(cudaMemset(*out, 0, n * sizeof(T)));
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_create(FdtdxPlan** out, int nx, int ny, int nz, int x_offset, int nx_global,
                                      double courant_number, double dt, int total_time_steps, int eps_tier,
                                      int mu_tier, int sigma_e_tier, int sigma_h_tier, double inv_mu_scalar,
                                      const int wrap[3], const float* const sB[3], const float* const sF[3],
                                      const float* const widths[3]) {
  if (!out || nx <= 0 || ny <= 0 || nz <= 0) return fail(FDTDX_EINVAL, "plan_create: bad dimensions");
  if (!(eps_tier == 1 || eps_tier == 3 || eps_tier == 9)) return fail(FDTDX_EINVAL, "eps_tier must be 1, 3 or 9");
  if (!(mu_tier == 0 || mu_tier == 1 || mu_tier == 3 || mu_tier == 9)) return fail(FDTDX_EINVAL, "mu_tier must be 0, 1, 3 or 9");
  FdtdxPlan* p = new FdtdxPlan();
  p->nx = nx; p->ny = ny; p->nz = nz; p->xoff = x_offset; p->nxg = nx_global; p->T = total_time_steps;
  p->courant = courant_number; p->dt = dt;
  p->eps_tier = eps_tier; p->mu_tier = mu_tier; p->sigE_tier = sigma_e_tier; p->sigH_tier = sigma_h_tier;
  p->inv_mu_scalar = inv_mu_scalar;
  for (int a = 0; a < 3; ++a) p->wrap[a] = wrap ? wrap[a] : 0;
  p->halo_lo_planned = (nx != nx_global) && (x_offset > 0 || p->wrap[0]);
  memset(p->slots, 0, sizeof(p->slots));
  p->metric = (sB != nullptr && sB[0] != nullptr);
  const int n[3] = {nx, ny, nz};
  const int ng[3] = {nx_global, ny, nz};
  for (int a = 0; a < 3; ++a) {
    p->d_sB[a] = p->d_sF[a] = p->d_w[a] = nullptr;
    if (p->metric) {
      int rc = to_device(p, sB[a], n[a], &p->d_sB[a]);
      if (rc) return rc;
      rc = to_device(p, sF[a], n[a], &p->d_sF[a]);
      if (rc) return rc;
    }
    if (widths && widths[a]) {
      int rc = to_device(p, widths[a], ng[a], &p->d_w[a]);
      if (rc) return rc;
    }
  }
  *out = p;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_destroy(FdtdxPlan* p) {
  if (p) for (auto& kv : p->ipc_open) cudaIpcCloseMemHandle(kv.second);
  if (!p) return FDTDX_OK;
  for (void* q : p->owned) cudaFree(q);
  for (int k = 0; k < 2; ++k) {
    if (p->adj_side[k]) cudaStreamDestroy(p->adj_side[k]);
    if (p->adj_join[k]) cudaEventDestroy(p->adj_join[k]);
  }
  if (p->adj_fork) cudaEventDestroy(p->adj_fork);
  delete p;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_pml_set_true_range(FdtdxPlan* p, int index, int lo, int hi) {
  if (!p || index < 0 || index >= (int)p->pmls.size()) return fail(FDTDX_EINVAL, "pml_set_true_range: bad slab index");
  PmlHost& h = p->pmls[index];
  if (lo < h.lo || hi > h.hi || hi <= lo) return fail(FDTDX_EINVAL, "pml_set_true_range: the true slab must lie inside the registered one");
  h.lo_true = lo;
  h.hi_true = hi;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_add_pml(FdtdxPlan* p, int axis, int direction, int lo, int hi, const float* a_E,
                                       const float* b_E, const float* inv_kappa_E, const float* a_H,
                                       const float* b_H, const float* inv_kappa_H, int kappa_is_one) {
  if (!p || axis < 0 || axis > 2 || hi <= lo) return fail(FDTDX_EINVAL, "add_pml: bad arguments");
  if ((int)p->pmls.size() >= FDTDX_MAX_PML) return fail(FDTDX_EINVAL, "add_pml: at most 6 slabs");
  const int ng[3] = {p->nxg, p->ny, p->nz};
  if (direction == 0 && lo != 0) return fail(FDTDX_EUNSUPPORTED, "add_pml: a '-' slab must start at index 0");
  if (direction == 1 && hi != ng[axis]) return fail(FDTDX_EUNSUPPORTED, "add_pml: a '+' slab must end at the domain edge");
  PmlHost h;
  h.axis = axis; h.dir = direction; h.lo = lo; h.hi = hi; h.kappa_one = kappa_is_one;
  h.lo_true = lo; h.hi_true = hi;
  const int L = hi - lo;
  h.aE.assign(a_E, a_E + L); h.bE.assign(b_E, b_E + L);
  h.aH.assign(a_H, a_H + L); h.bH.assign(b_H, b_H + L);
  h.kE.resize(L); h.kH.resize(L);
  for (int i = 0; i < L; ++i) {
    h.kE[i] = inv_kappa_E[i] - 1.0f;
    h.kH[i] = inv_kappa_H[i] - 1.0f;
  }
  p->pmls.push_back(h);
  p->finalized = false;
  return (int)p->pmls.size() - 1;
}

extern "C" int fdtdx_b200_plan_add_wall(FdtdxPlan* p, int kind, int axis, const int lo[3], const int hi[3]) {
  if (!p || (int)p->walls.size() >= FDTDX_MAX_WALL) return fail(FDTDX_EINVAL, "add_wall: too many walls");
  WallDev w;
  w.kind = kind; w.axis = axis;
  for (int a = 0; a < 3; ++a) { w.lo[a] = lo[a]; w.hi[a] = hi[a]; }
  w.lo[0] -= p->xoff; w.hi[0] -= p->xoff;
  p->walls.push_back(w);
  p->finalized = false;
  return (int)p->walls.size() - 1;
}

static int fill_profile(FdtdxPlan* p, SrcDev& d, int profile_kind, const double* q, const float* signal, int signal_len) {
  d.profile_kind = profile_kind;
  for (int i = 0; i < 6; ++i) d.p[i] = 0.f;
  if (profile_kind == FDTDX_PROFILE_CW) {
    d.p[0] = (float)q[0]; d.p[1] = (float)q[1]; d.p[2] = (float)q[2]; d.p[3] = (float)q[3];
    d.p[4] = (float)(2.0 * M_PI);
  } else if (profile_kind == FDTDX_PROFILE_PULSE) {
    d.p[0] = (float)q[0]; d.p[1] = (float)q[1]; d.p[2] = (float)q[2]; d.p[3] = (float)q[3];
    d.p[5] = (float)q[4];
  } else {
    d.p[0] = (float)q[0]; d.p[1] = (float)q[1]; d.p[2] = (float)q[2]; d.p[3] = (float)q[3];
  }
  d.signal = nullptr; d.signal_len = signal_len;
  if (signal && signal_len > 0) {
    float* ds;
    int rc = to_device(p, signal, signal_len, &ds);
    if (rc) return rc;
    d.signal = ds;
  }
  return FDTDX_OK;
}

static int fill_switch(FdtdxPlan* p, SrcHost& s, const uint8_t* on, const float* t_adj) {
  s.d.on = nullptr; s.d.t_adj = nullptr;
  if (on) {
    s.on.assign(on, on + p->T);
    uint8_t* d_on; float* d_t;
    int rc = to_device(p, on, p->T, &d_on);
    if (rc) return rc;
    rc = to_device(p, t_adj, p->T, &d_t);
    if (rc) return rc;
    s.d.on = d_on; s.d.t_adj = d_t;
  }
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_add_plane_source(FdtdxPlan* p, const int lo[3], const int hi[3], int normal_axis,
                                                int sign, const float* E_inc, const float* H_inc,
                                                const float* toff_E, const float* toff_H, int profile_kind,
                                                const double params[8], const float* signal, int signal_len,
                                                double static_amplitude, double cE, double cH, const uint8_t* on,
                                                const float* t_adj, const float* h_filter, int h_filter_len) {
  if (!p || (int)p->srcs.size() >= FDTDX_MAX_SRC) return fail(FDTDX_EINVAL, "add_plane_source: too many sources");
  SrcHost s;
  memset(&s.d, 0, sizeof(SrcDev));
  s.d.kind = 0;
  // clip the plane to this rank's x-slab; the face arrays are sliced accordingly by the caller
  for (int a = 0; a < 3; ++a) { s.d.lo[a] = lo[a]; s.d.hi[a] = hi[a]; }
  s.d.lo[0] -= p->xoff; s.d.hi[0] -= p->xoff;
  s.d.normal_axis = normal_axis;
  s.d.sign = (float)sign;
  s.d.static_amp = (float)static_amplitude;
  s.d.cE = (float)cE; s.d.cH = (float)cH;
  const size_t fn = (size_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
  float *dE, *dH, *dtE, *dtH;
  int rc;
  if ((rc = to_device(p, E_inc, 3 * fn, &dE))) return rc;
  if ((rc = to_device(p, H_inc, 3 * fn, &dH))) return rc;
  if ((rc = to_device(p, toff_E, 3 * fn, &dtE))) return rc;
  if ((rc = to_device(p, toff_H, 3 * fn, &dtH))) return rc;
  s.d.Einc = dE; s.d.Hinc = dH; s.d.toffE = dtE; s.d.toffH = dtH;
  if ((rc = fill_profile(p, s.d, profile_kind, params, signal, signal_len))) return rc;
  if ((rc = fill_switch(p, s, on, t_adj))) return rc;
  s.d.hfilter = nullptr; s.d.hfilter_len = 0;
  if (h_filter && h_filter_len > 1) {
    float* dh;
    if ((rc = to_device(p, h_filter, h_filter_len, &dh))) return rc;
    s.d.hfilter = dh; s.d.hfilter_len = h_filter_len;
  }
  p->srcs.push_back(s);
  p->finalized = false;
  return (int)p->srcs.size() - 1;
}

extern "C" int fdtdx_b200_plan_source_set_quadrature(FdtdxPlan* p, int source_index, const float* E_inc_imag,
                                                      const float* H_inc_imag, double quadrature_phase) {
  if (!p || source_index < 0 || source_index >= (int)p->srcs.size() || !E_inc_imag || !H_inc_imag)
    return fail(FDTDX_EINVAL, "source_set_quadrature: bad arguments");
  SrcDev& d = p->srcs[source_index].d;
  if (d.kind != 0) return fail(FDTDX_EINVAL, "source_set_quadrature: not a plane source");
  const size_t fn = (size_t)(d.hi[0] - d.lo[0]) * (d.hi[1] - d.lo[1]) * (d.hi[2] - d.lo[2]);
  float *dE, *dH;
  int rc;
  if ((rc = to_device(p, E_inc_imag, 3 * fn, &dE))) return rc;
  if ((rc = to_device(p, H_inc_imag, 3 * fn, &dH))) return rc;
  d.EincI = dE; d.HincI = dH; d.pq = (float)quadrature_phase;
  p->finalized = false;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_add_dipole(FdtdxPlan* p, const int cell[3], int polarization, int electric,
                                          double scale, int profile_kind, const double params[8],
                                          const float* signal, int signal_len, const uint8_t* on,
                                          const float* t_adj) {
  if (!p || (int)p->srcs.size() >= FDTDX_MAX_SRC) return fail(FDTDX_EINVAL, "add_dipole: too many sources");
  SrcHost s;
  memset(&s.d, 0, sizeof(SrcDev));
  s.d.kind = 1;
  for (int a = 0; a < 3; ++a) { s.d.lo[a] = cell[a]; s.d.hi[a] = cell[a] + 1; }
  s.d.lo[0] -= p->xoff; s.d.hi[0] -= p->xoff;
  s.d.pol = polarization; s.d.electric = electric; s.d.dip_scale = (float)scale;
  int rc;
  if ((rc = fill_profile(p, s.d, profile_kind, params, signal, signal_len))) return rc;
  if ((rc = fill_switch(p, s, on, t_adj))) return rc;
  p->srcs.push_back(s);
  p->finalized = false;
  return (int)p->srcs.size() - 1;
}

extern "C" int fdtdx_b200_plan_add_detector(FdtdxPlan* p, int kind, const int lo[3], const int hi[3], int flags,
                                            int comp_mask, int aux, const uint8_t* on, const int32_t* arr_idx,
                                            const float* weights, int n_freq, const float* phasor_table,
                                            const float* window, double scale, const int slice_idx[3]) {
  if (!p || !on || !arr_idx) return fail(FDTDX_EINVAL, "add_detector: missing tables");
  DetHost h;
  memset(&h.d, 0, sizeof(DetDev));
  DetDev& d = h.d;
  d.kind = kind; d.flags = flags; d.comp_mask = comp_mask; d.aux = aux;
  for (int a = 0; a < 3; ++a) { d.lo[a] = lo[a]; d.hi[a] = hi[a]; d.slice_idx[a] = slice_idx ? slice_idx[a] : 0; }
  d.lo[0] -= p->xoff; d.hi[0] -= p->xoff;  // local x coordinates of this rank's slab
  if (p->nx != p->nxg) {
    // x-sharded: the region and its co-location stencil (x-1) must lie inside this slab; detectors
    // straddling a slab edge would need an extra E-plane exchange - not built yet.
    // the caller passes this rank's part of the region; an exact-interpolation detector that starts on
    // plane 0 of a slab with a lower neighbour reads that neighbour's last plane from the DET_XLO_* buffers
    // (on the first slab of a non-periodic axis plane -1 is the zero halo, as on an unsharded grid)
    if (d.lo[0] < 0 || d.hi[0] > p->nx)
      return fail(FDTDX_EUNSUPPORTED, "detector region must be clipped to this rank's x-slab");
  }
  d.ncomp = 0;
  for (int c = 0; c < 6; ++c) d.ncomp += (comp_mask >> c) & 1;
  const size_t n = (size_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
  h.on.assign(on, on + p->T);
  int rc;
  uint8_t* d_on; int32_t* d_idx;
  if ((rc = to_device(p, on, p->T, &d_on))) return rc;
  if ((rc = to_device(p, arr_idx, p->T, &d_idx))) return rc;
  d.on = d_on; d.arr_idx = d_idx;
  d.weights = nullptr; d.wsum = 1.0f;
  const bool keep_all = (kind == FDTDX_DET_POYNTING) && (flags & DET_KEEP_ALL);
  if (weights) {
    const size_t wn = keep_all ? 3 * n : n;
    float* dw;
    if ((rc = to_device(p, weights, wn, &dw))) return rc;
    d.weights = dw;
    // jnp.sum(weights) in float32 (detector.py:108); pairwise order as numpy does it is close
    // enough for a normalisation constant - computed in double and rounded once.
    double acc = 0.0;
    for (size_t i = 0; i < n; ++i) acc += (double)weights[i];
    d.wsum = (float)acc;
  }
  d.nf = n_freq; d.scale = (float)scale;
  if (kind == FDTDX_DET_PHASOR) {
    if (!phasor_table || !window) return fail(FDTDX_EINVAL, "add_detector: phasor needs table and window");
    float* dt; float* dw;
    if ((rc = to_device(p, phasor_table, (size_t)2 * p->T * n_freq, &dt))) return rc;
    if ((rc = to_device(p, window, p->T, &dw))) return rc;
    d.ph_table = reinterpret_cast<const float2*>(dt);
    d.window = dw;
  }
  // large regions: aligned H_prev rows + row-marching kernels; needs 16-byte rows (Nz % 4 == 0)
  const size_t Ngrid = (size_t)p->nx * p->ny * p->nz;
  h.volume = (flags & DET_VOLUME) && (p->nz % 4 == 0) && (3 * Ngrid * sizeof(float) <= ((size_t)16 << 30));
  d.flags &= ~DET_VOLUME;  // set per launch by sync_dets once the bound buffers are known to be aligned
  d.hz0 = lo[2] & ~3;
  d.hrow = ((hi[2] + 1 - d.hz0) + 3) & ~3;
  if (flags & DET_EXACT) {
    const size_t hn = (size_t)3 * (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
    if ((rc = to_device<float>(p, nullptr, hn, &d.hprev))) return rc;
    if (h.volume) {  // one (3,Nx,Ny,Nz) H_prev scratch for all row-marching detectors
      if (!p->d_hprev_full && (rc = to_device<float>(p, nullptr, 3 * Ngrid, &p->d_hprev_full))) return rc;
      d.hprev_full = p->d_hprev_full;
    }
  }
  if (h.volume && kind == FDTDX_DET_ENERGY && (flags & DET_SLICES) && (flags & DET_SLICE_MEAN)) {
    // partial sums of the three slice means: over z tiles, over y tiles, over x chunks (>= 4 planes each)
    const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    d.npart[0] = (hi[2] - d.hz0 + DETV_TZ - 1) / DETV_TZ;
    d.npart[1] = (ey + DETV_ROWS - 1) / DETV_ROWS;
    d.npart[2] = (ex + 3) / 4;
    if ((rc = to_device<float>(p, nullptr, (size_t)d.npart[0] * ex * ey, &d.part[0]))) return rc;
    if ((rc = to_device<float>(p, nullptr, (size_t)d.npart[1] * ex * ez, &d.part[1]))) return rc;
    if ((rc = to_device<float>(p, nullptr, (size_t)d.npart[2] * ey * ez, &d.part[2]))) return rc;
  }
  h.nvals = 0;
  const bool staged = (flags & DET_REDUCE) || ((flags & DET_SLICES) && (flags & DET_SLICE_MEAN));
  if (staged) {
    h.nvals = (kind == FDTDX_DET_FIELD || kind == FDTDX_DET_PHASOR) ? d.ncomp : ((keep_all && !(flags & DET_CLOSED)) ? 3 : 1);
    if ((rc = to_device<float>(p, nullptr, (size_t)h.nvals * n, &d.scratch))) return rc;
  }
  p->dets.push_back(h);
  p->dets_dirty = true;
  return (int)p->dets.size() - 1;
}

extern "C" int fdtdx_b200_plan_detector_set_wsum(FdtdxPlan* p, int detector_index, double weight_sum) {
  if (!p || detector_index < 0 || detector_index >= (int)p->dets.size()) return fail(FDTDX_EINVAL, "detector_set_wsum: bad detector index");
  p->dets[detector_index].d.wsum = (float)weight_sum;
  p->dets_dirty = true;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_set_recorder(FdtdxPlan* p, int dtype, int n_slots, const int32_t* slot_of_time,
                                            const int32_t* replay_a, const int32_t* replay_b, const float* replay_w) {
  if (!p || !slot_of_time) return fail(FDTDX_EINVAL, "set_recorder: missing tables");
  p->has_rec = true; p->rec_dtype = dtype; p->rec_slots = n_slots;
  p->slot_of_time.assign(slot_of_time, slot_of_time + p->T);
  p->replay_a.assign(replay_a, replay_a + p->T);
  p->replay_b.assign(replay_b, replay_b + p->T);
  p->replay_w.assign(replay_w, replay_w + p->T);
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_plan_set_dispersion(FdtdxPlan* p, int n_poles, int coeff_tier, int has_c4) {
  if (!p || n_poles < 0 || !(coeff_tier == 1 || coeff_tier == 3)) return fail(FDTDX_EINVAL, "set_dispersion: bad arguments");
  p->n_poles = n_poles; p->coeff_tier = coeff_tier; p->has_c4 = has_c4;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_halo_bind(FdtdxPlan* p, int has_lo, int has_hi) {
  if (!p) return fail(FDTDX_EINVAL, "halo_bind: null plan");
  p->halo_lo = has_lo; p->halo_hi = has_hi;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_set_z_padding(FdtdxPlan* p, int pad_cells) {
  if (!p || pad_cells < 0 || pad_cells >= p->nz) return fail(FDTDX_EINVAL, "set_z_padding: bad argument");
  p->zpad = pad_cells;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_set_symmetry(FdtdxPlan* p, const int symmetric_axes[3], const int electric_wall_axes[3]) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  for (int a = 0; a < 3; ++a) {
    p->sym[a] = (symmetric_axes && symmetric_axes[a]) ? 1 : 0;
    p->mirror[a] = (p->sym[a] && electric_wall_axes && electric_wall_axes[a]) ? 1 : 0;
  }
  if ((p->sym[0] | p->sym[1] | p->sym[2]) && (p->eps_tier == 9 || p->mu_tier == 9 || p->sigE_tier == 9 || p->sigH_tier == 9 || p->nx != p->nxg))
    return fail(FDTDX_EUNSUPPORTED, "config.symmetry with full-tensor media or on x-sharded plans");
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_set_bloch(FdtdxPlan* p, int enable, const double cos_kL[3], const double sin_kL[3]) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  p->bloch = enable != 0;
  for (int a = 0; a < 3; ++a) {
    p->bloch_c[a] = (enable && cos_kL) ? (float)cos_kL[a] : 1.0f;
    p->bloch_s[a] = (enable && sin_kL) ? (float)sin_kL[a] : 0.0f;
  }
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_bind(FdtdxPlan* p, int slot, int index, void* ptr) {
  if (!p || slot < 0 || slot >= FDTDX_SLOT_COUNT || index < 0 || index >= MAX_IDX)
    return fail(FDTDX_EINVAL, "bind: bad slot/index");
  p->slots[slot][index] = ptr;
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_get_parity(FdtdxPlan* p, int* pp, int* ep, int* hp) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  if (pp) *pp = p->p_parity;
  if (ep) *ep = p->e_parity;
  if (hp) *hp = p->h_parity;
  return FDTDX_OK;
}
extern "C" int fdtdx_b200_set_parity(FdtdxPlan* p, int pp, int ep, int hp) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  p->p_parity = pp & 1; p->e_parity = ep & 1; p->h_parity = hp & 1;
  return FDTDX_OK;
}
extern "C" long long fdtdx_b200_launch_count(FdtdxPlan* p) { return p ? p->launches : 0; }
extern "C" int fdtdx_b200_set_tma(FdtdxPlan* p, int enable, int xchunk_tma) {
  if (!p) return fail(FDTDX_EINVAL, "set_tma: null plan");
  p->use_tma = enable;
  p->xchunk_tma = xchunk_tma;
  return FDTDX_OK;
}
extern "C" int fdtdx_b200_set_tuning(FdtdxPlan* p, int xchunk, int rows) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  p->xchunk = xchunk;
  if (rows > 0) p->rows = std::min(rows, 8);
  return FDTDX_OK;
}

// ------------------------------------------------------------------------------------------------
static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

static int finalize(FdtdxPlan* p) {
  if (p->finalized) return FDTDX_OK;
  const int n[3] = {p->nx, p->ny, p->nz};
  for (int a = 0; a < 3; ++a) {
    AxisPmlDev& A = p->axis[a];
    memset(&A, 0, sizeof(A));
    A.lo_len = 0; A.hi_start = n[a]; A.hi_len = 0; A.kappa_one = 1;
    std::vector<float> t[6];
    for (auto& v : t) v.assign(n[a], 0.0f);
    const int off = (a == 0) ? p->xoff : 0;
    for (const PmlHost& h : p->pmls) {
      if (h.axis != a) continue;
      const int llo = std::max(h.lo - off, 0), lhi = std::min(h.hi - off, n[a]);
      if (lhi <= llo) continue;
      if (h.dir == 0) {
        if (llo != 0) return fail(FDTDX_EUNSUPPORTED, "x-PML slab split across ranks away from the slab start");
        A.lo_len = lhi;
      } else {
        if (lhi != n[a]) return fail(FDTDX_EUNSUPPORTED, "x-PML slab split across ranks away from the slab end");
        A.hi_start = llo; A.hi_len = lhi - llo;
      }
      if (!h.kappa_one) A.kappa_one = 0;
      for (int i = llo; i < lhi; ++i) {
        const int q = i + off - h.lo;
        t[0][i] = h.aE[q]; t[1][i] = h.bE[q]; t[2][i] = h.kE[q];
        t[3][i] = h.aH[q]; t[4][i] = h.bH[q]; t[5][i] = h.kH[q];
      }
    }
    if (A.lo_len > A.hi_start)
      return fail(FDTDX_EUNSUPPORTED, "the '-' and '+' CPML slabs of one axis overlap (grid thinner than twice the PML thickness)");
    float* d[6];
    for (int q = 0; q < 6; ++q) {
      int rc = to_device(p, t[q].data(), (size_t)n[a], &d[q]);
      if (rc) return rc;
    }
    A.aE = d[0]; A.bE = d[1]; A.kE = d[2]; A.aH = d[3]; A.bH = d[4]; A.kH = d[5];
  }
  if (!p->srcs.empty()) {
    std::vector<SrcDev> hs;
    for (auto& s : p->srcs) hs.push_back(s.d);
    int rc = to_device(p, hs.data(), hs.size(), &p->d_srcs);
    if (rc) return rc;
  }
  if (!p->walls.empty()) {
    int rc = to_device(p, p->walls.data(), p->walls.size(), &p->d_walls);
    if (rc) return rc;
  }
  p->finalized = true;
  return FDTDX_OK;
}

static int make_params(FdtdxPlan* p, StepParams& P, int simulate) {
  memset(&P, 0, sizeof(P));
  P.nx = p->nx; P.ny = p->ny; P.nz = p->nz;
  const long long N = (long long)p->nx * p->ny * p->nz;
  for (int a = 0; a < 3; ++a) P.wrap[a] = p->wrap[a];
  for (int a = 0; a < 3; ++a) P.sym[a] = p->sym[a];
  P.x_lo_mode = p->halo_lo ? 2 : ((p->wrap[0] && !p->sym[0]) ? 1 : 0);
  P.x_hi_mode = p->halo_hi ? 2 : (p->wrap[0] ? 1 : 0);
  P.cour = (float)p->courant; P.eta0 = (float)kEta0; P.inv_mu_scalar = (float)p->inv_mu_scalar; P.dt = (float)p->dt;
  // full-tensor tiers ping-pong their field between the primary and the ALT buffer
  P.E = (float*)p->slots[p->e_parity ? FDTDX_SLOT_E_ALT : FDTDX_SLOT_E][0];
  P.H = (float*)p->slots[p->h_parity ? FDTDX_SLOT_H_ALT : FDTDX_SLOT_H][0];
  P.eps = (const float*)p->slots[FDTDX_SLOT_INV_EPS][0];
  P.mu = (const float*)p->slots[FDTDX_SLOT_INV_MU][0];
  P.sigE = (const float*)p->slots[FDTDX_SLOT_SIGMA_E][0];
  P.sigH = (const float*)p->slots[FDTDX_SLOT_SIGMA_H][0];
  if (!P.E || !P.H || !P.eps) return fail(FDTDX_EUNBOUND, "E, H and INV_EPS must be bound");
  if (p->mu_tier > 0 && !P.mu) return fail(FDTDX_EUNBOUND, "INV_MU must be bound for mu_tier > 0");
  if (p->sigE_tier > 0 && !P.sigE) return fail(FDTDX_EUNBOUND, "SIGMA_E must be bound");
  if (p->sigH_tier > 0 && !P.sigH) return fail(FDTDX_EUNBOUND, "SIGMA_H must be bound");
  P.eps_cs = (p->eps_tier == 1) ? 0 : N;
  P.mu_cs = (p->mu_tier <= 1) ? 0 : N;
  P.sigE_cs = (p->sigE_tier <= 1) ? 0 : N;
  P.sigH_cs = (p->sigH_tier <= 1) ? 0 : N;
  for (int a = 0; a < 3; ++a) { P.sB[a] = p->d_sB[a]; P.sF[a] = p->d_sF[a]; }
  for (int a = 0; a < 3; ++a) P.pml[a] = p->axis[a];
  for (size_t q = 0; q < p->pmls.size(); ++q) {
    const PmlHost& h = p->pmls[q];
    const int off = (h.axis == 0) ? p->xoff : 0;
    const int n[3] = {p->nx, p->ny, p->nz};
    if (std::min(h.hi - off, n[h.axis]) <= std::max(h.lo - off, 0)) continue;  // slab not on this rank
    for (int w = 0; w < 2; ++w) {
      float* pe = (float*)p->slots[FDTDX_SLOT_PSI_E][2 * q + w];
      float* ph = (float*)p->slots[FDTDX_SLOT_PSI_H][2 * q + w];
      if (!pe || !ph) return fail(FDTDX_EUNBOUND, "PSI_E / PSI_H must be bound for every PML slab");
      P.pml[h.axis].psiE[h.dir][w] = pe;
      P.pml[h.axis].psiH[h.dir][w] = ph;
    }
  }
  {
    // z slabs can move psi as float2 halves when slab thickness, slab start and nz are even/4-aligned
    // and no lane (4 cells) touches both slabs
    AxisPmlDev& Z = P.pml[2];
    bool ok = (p->nz % 4 == 0) && (Z.lo_len % 2 == 0) && (Z.hi_len % 2 == 0) && (Z.hi_start % 2 == 0) &&
              (Z.lo_len + 4 <= Z.hi_start || Z.lo_len == 0 || Z.hi_len == 0);
    for (int s = 0; s < 2; ++s)
      for (int w = 0; w < 2; ++w) {
        if (Z.psiE[s][w] && (reinterpret_cast<uintptr_t>(Z.psiE[s][w]) & 7u)) ok = false;
        if (Z.psiH[s][w] && (reinterpret_cast<uintptr_t>(Z.psiH[s][w]) & 7u)) ok = false;
      }
    Z.vec_ok = ok ? 1 : 0;
  }
  P.simulate = simulate;
  P.psi_store = 1;
  P.p_store = 1;
  P.n_walls = (int)p->walls.size();
  P.walls = p->d_walls;
  P.n_src = (int)p->srcs.size();
  P.src = p->d_srcs;
  P.src_x0 = p->nx; P.src_x1 = 0;
  for (int s = 0; s < P.n_src; ++s) {
    const SrcDev& d = p->srcs[s].d;
    for (int a = 0; a < 3; ++a) { P.src_lo[s][a] = d.lo[a]; P.src_hi[s][a] = d.hi[a]; }
    P.src_x0 = std::min(P.src_x0, d.lo[0]);
    P.src_x1 = std::max(P.src_x1, d.hi[0]);
  }
  // x-normal planes and dipoles are injected inside the half-step kernels; sources that cross many x
  // planes (y / z-normal planes, volumes) by src_apply_kernel (source_kernels.cuh)
  P.src_inline = (P.src_x1 - P.src_x0 <= 2) ? 1 : 0;
  {
    const char* e = getenv("FDTDX_B200_SRC_INLINE");
    if (e) P.src_inline = (e[0] != '0');
  }
  for (int k = 0; k < 2; ++k) { P.wall_x0[k] = p->nx; P.wall_x1[k] = 0; }
  for (int w = 0; w < P.n_walls; ++w) {
    const WallDev& W = p->walls[w];
    P.wallp[w] = W;
    P.wall_x0[W.kind] = std::min(P.wall_x0[W.kind], W.lo[0]);
    P.wall_x1[W.kind] = std::max(P.wall_x1[W.kind], W.hi[0]);
  }
  P.n_poles = p->n_poles; P.has_c4 = p->has_c4;
  if (p->n_poles > 0) {
    float* A = (float*)p->slots[FDTDX_SLOT_P_A][0];
    float* B = (float*)p->slots[FDTDX_SLOT_P_B][0];
    if (!A || !B) return fail(FDTDX_EUNBOUND, "P_A / P_B must be bound for dispersive runs");
    P.P_cur = p->p_parity ? B : A;
    P.P_new = p->p_parity ? A : B;
    P.c1 = (const float*)p->slots[FDTDX_SLOT_C1][0];
    P.c2 = (const float*)p->slots[FDTDX_SLOT_C2][0];
    P.c3 = (const float*)p->slots[FDTDX_SLOT_C3][0];
    P.c4 = (const float*)p->slots[FDTDX_SLOT_C4][0];
    if (!P.c1 || !P.c2 || !P.c3 || (p->has_c4 && !P.c4)) return fail(FDTDX_EUNBOUND, "dispersive coefficients must be bound");
    P.c_cs = (p->coeff_tier == 1) ? 0 : N;
  }
  if (p->bloch) {
    P.bE = (const float*)p->slots[FDTDX_SLOT_BLOCH_E][0];
    P.bH = (const float*)p->slots[FDTDX_SLOT_BLOCH_H][0];
    if (!P.bE || !P.bH) return fail(FDTDX_EUNBOUND, "BLOCH_E / BLOCH_H (the partner system's fields) must be bound");
    for (int a = 0; a < 3; ++a) { P.bc[a] = p->bloch_c[a]; P.bs[a] = p->bloch_s[a]; }
  }
  P.haloH = (const float*)p->slots[FDTDX_SLOT_HALO_H_LO][0];
  P.haloE = (const float*)p->slots[FDTDX_SLOT_HALO_E_HI][0];
  P.haloH_cs = P.haloE_cs = (long long)p->ny * p->nz;
  if (p->peer_mode) {
    // read the neighbour's boundary plane in place: Hy of its last plane / Ey of its first plane
    const long long plane = (long long)p->ny * p->nz;
    if (p->halo_lo) {
      const long long Np = plane * p->peer[0].nx;
      P.haloH = p->peer[0].field + Np + (long long)(p->peer[0].nx - 1) * plane;
      P.haloH_cs = Np;
    }
    if (p->halo_hi) {
      const long long Np = plane * p->peer[1].nx;
      P.haloE = p->peer[1].field + Np;
      P.haloE_cs = Np;
    }
  }
  if (p->halo_lo && !P.haloH) return fail(FDTDX_EUNBOUND, "HALO_H_LO must be bound");
  if (p->halo_hi && !P.haloE) return fail(FDTDX_EUNBOUND, "HALO_E_HI must be bound");
  int xc = p->xchunk;
  if (xc <= 0) {
    // enough CTAs for >= ~6 waves of 148 SMs x 8 resident CTAs, but chunks long enough that the
    // one-plane re-read at each chunk start stays below ~3 % of the traffic
    const long long tiles = (long long)((p->nz + 127) / 128) * ((p->ny + p->rows - 1) / p->rows);
    long long want = (148LL * 2 * 16 + tiles - 1) / tiles;  // >= ~16 waves at 2 resident CTAs per SM
    // small grids (about a wave of CTAs): shorter chunks balance the load better than they cost in re-read planes
    // (C3b 120^3, periodic x / y: 112 us/step at 16 planes per CTA, 95 at 8 - scripts/c3b_perf.py)
    xc = (int)std::max(8LL, std::min<long long>(32, p->nx / std::max(1LL, want)));
  }
  P.xchunk = std::min(xc, p->nx);
  P.x_begin = 0;
  P.x_end = p->nx;
  return FDTDX_OK;
}

static bool can_vec4(const FdtdxPlan* p, const StepParams& P) {
  if (p->nz % 4 != 0) return false;
  {
    // the 4-cells-per-lane CPML paths assume a lane never touches both z slabs (tiny grids only)
    const AxisPmlDev& Z = P.pml[2];
    if (Z.lo_len > 0 && Z.hi_len > 0 && (Z.lo_len - 1) / 4 == Z.hi_start / 4) return false;
  }
  const void* ptrs[] = {P.E, P.H, P.eps, P.mu, P.haloH, P.haloE, P.sigE, P.sigH, P.P_cur, P.P_new, P.c1, P.c2, P.c3, P.c4};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return false;
  for (int a = 0; a < 2; ++a)
    for (int s = 0; s < 2; ++s)
      for (int w = 0; w < 2; ++w) {
        if (P.pml[a].psiE[s][w] && !aligned16(P.pml[a].psiE[s][w])) return false;
        if (P.pml[a].psiH[s][w] && !aligned16(P.pml[a].psiH[s][w])) return false;
      }
  return true;
}

// kernel dispatchers live in yee_E*.cu / yee_H*.cu (separate translation units, built in parallel)
void fdtdx_dispatch_E4(const StepParams& P, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st);
void fdtdx_dispatch_E1(const StepParams& P, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g, dim3 b, cudaStream_t st);
void fdtdx_dispatch_H4(const StepParams& P, int t, int mu_tier, int pm, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st);
void fdtdx_dispatch_H1(const StepParams& P, int t, int mu_tier, int pm, bool rev, bool sig, bool met, dim3 g, dim3 b, cudaStream_t st);
void fdtdx_dispatch_E4_konly(const StepParams& P, int t, int pm, bool rev, bool met, dim3 g, dim3 b, cudaStream_t st);
void fdtdx_dispatch_H4_konly(const StepParams& P, int t, int pm, bool rev, bool met, dim3 g, dim3 b, cudaStream_t st);
struct alignas(64) TmaSet {
  CUtensorMap fld_halo, fld_plain, mat_plain, xhalo;
};
cudaError_t fdtdx_dispatch_E4_tma(const StepParams& P, const TmaSet& M, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g,
                                  cudaStream_t st);
cudaError_t fdtdx_dispatch_H4_tma(const StepParams& P, const TmaSet& M, int t, int mt, int pm, bool rev, bool sig, bool met, dim3 g, cudaStream_t st);
cudaError_t fdtdx_dispatch_E4_tma_konly(const StepParams& P, const TmaSet& M, int t, int pm, bool rev, bool met, dim3 g, cudaStream_t st);
cudaError_t fdtdx_dispatch_H4_tma_konly(const StepParams& P, const TmaSet& M, int t, int pm, bool rev, bool met, dim3 g, cudaStream_t st);
// 64-cell tile rows, two rows per warp: thin grids (Nz <= 64)
cudaError_t fdtdx_dispatch_E4_tma64(const StepParams& P, const TmaSet& M, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g,
                                    cudaStream_t st);
cudaError_t fdtdx_dispatch_H4_tma64(const StepParams& P, const TmaSet& M, int t, int mt, int pm, bool rev, bool sig, bool met, dim3 g, cudaStream_t st);
// flat tiles: any other row length up to 124 cells (the CTA's threads laid over (row, z quad) pairs)
cudaError_t fdtdx_dispatch_E4_tmaf(const StepParams& P, const TmaSet& M, int t, int tier, int pm, bool rev, bool sig, bool ade, bool met, dim3 g,
                                   cudaStream_t st);
cudaError_t fdtdx_dispatch_H4_tmaf(const StepParams& P, const TmaSet& M, int t, int mt, int pm, bool rev, bool sig, bool met, dim3 g, cudaStream_t st);
static int tma_tz(const FdtdxPlan* p) {
  const char* e = getenv("FDTDX_B200_TMA_TZ");
  if (e && (atoi(e) == 64 || atoi(e) == 128)) return atoi(e);
  return p->nz <= 64 ? 64 : 128;
}
// z quads per row when the half-steps run with flat tiles (0: the 64- / 128-cell tile rows).  FDTDX_B200_TMA_FLAT=0
// or an explicit FDTDX_B200_TMA_TZ turn it off.
static int tma_flat_lz(const FdtdxPlan* p) {
  const char* e = getenv("FDTDX_B200_TMA_FLAT");
  if (e && e[0] == '0') return 0;
  const char* tz = getenv("FDTDX_B200_TMA_TZ");
  if (tz && (atoi(tz) == 64 || atoi(tz) == 128)) return 0;
  if (p->nz % 4 != 0) return 0;
  // Rows of up to 124 cells are one flat tile; longer rows that are not a multiple of 128 cells are cut into
  // nt equal flat tiles (Nz = 136: two tiles of 17 quads instead of 128 + 8 cells) when that keeps clearly more
  // lanes busy than 128-cell tile rows would.
  const int q = p->nz / 4;                       // z quads per row
  const int nt = (q + 30) / 31;                  // tiles of at most 31 quads
  const int lz = (q + nt - 1) / nt;
  if (lz < 5 || lz > 31) return 0;
  if (nt == 1) return lz != 16 ? lz : 0;         // 64-cell rows: the two-rows-per-warp instantiation
  const int rt = (FDTDX_TMA_R * 32) / lz;
  const double flat_eff = (double)q / (double)(lz * nt) * ((double)(lz * rt) / (FDTDX_TMA_R * 32.0));
  const double std_eff = (double)p->nz / (128.0 * ((p->nz + 127) / 128));
  return flat_eff > std_eff + 0.05 ? lz : 0;
}
// tile row length / rows per CTA tile of the staged half-steps
static void tma_tile(const FdtdxPlan* p, int lz, int* tz, int* rt) {
  if (lz > 0) { *tz = 4 * lz; *rt = (FDTDX_TMA_R * 32) / lz; }
  else { *tz = tma_tz(p); *rt = FDTDX_TMA_R * (128 / *tz); }
}
// warps of one x chunk that own rows (the in-kernel peer ordering counts their arrivals)
static int tma_active_warps(const FdtdxPlan* p, int lz, int gx) {
  int tz, rt;
  tma_tile(p, lz, &tz, &rt);
  const int lanes = tz / 4;
  int n = 0;
  for (int j0 = 0; j0 < p->ny; j0 += rt) n += std::min(FDTDX_TMA_R, (std::min(rt, p->ny - j0) * lanes + 31) / 32);
  return gx * n;
}

// 0: no CPML slab on this rank; 1: scalar z-slab accesses; 2: 128-bit z-slab accesses
static int pml_mode(const FdtdxPlan* p, const StepParams& P) {
  bool any = false;
  const int n[3] = {p->nx, p->ny, p->nz};
  for (int a = 0; a < 3; ++a)
    if (P.pml[a].lo_len > 0 || P.pml[a].hi_start < n[a]) any = true;
  if (!any) return 0;
  return P.pml[2].vec_ok ? 2 : 1;
}

// ---- TMA-staged path (yee_tma.cuh) --------------------------------------------------------------
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)q;
    cudaGetLastError();
  }
  return fn;
}

// kind 0: halo box (TZ+4, R+1), kind 1: plain box (TZ, R).  Arrays are (C, nx, ny, nz) float32.
static int get_tmap(FdtdxPlan* p, const void* base, int comps, int nx, int kind, CUtensorMap* out, int lz = 0) {
  int tz, rt;
  tma_tile(p, lz, &tz, &rt);
  auto key = std::make_tuple(base, comps * 4 + (nx == p->nx ? 0 : 1) + 64 * lz, kind);
  auto it = p->tmaps.find(key);
  if (it != p->tmaps.end()) { *out = it->second; return FDTDX_OK; }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(FDTDX_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)p->nz, (cuuint64_t)p->ny, (cuuint64_t)nx, (cuuint64_t)comps};
  const cuuint64_t strides[3] = {(cuuint64_t)p->nz * 4, (cuuint64_t)p->nz * p->ny * 4, (cuuint64_t)p->nz * p->ny * nx * 4};
  const cuuint32_t box[4] = {(cuuint32_t)(kind == 0 ? tz + 4 : tz), (cuuint32_t)(kind == 0 ? rt + 1 : rt), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FDTDX_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  p->tmaps[key] = m;
  *out = m;
  return FDTDX_OK;
}

// (2, ny, nz)-shaped view with an arbitrary component stride (packed staging buffer or a peer field array)
static int get_xhalo_tmap(FdtdxPlan* p, const void* base, long long comp_stride, CUtensorMap* out, int lz = 0) {
  auto key = std::make_tuple(base, (int)(comp_stride % 2147483647LL), 2 + 4 * lz);
  auto it = p->tmaps.find(key);
  if (it != p->tmaps.end()) { *out = it->second; return FDTDX_OK; }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(FDTDX_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)p->nz, (cuuint64_t)p->ny, 2, 1};
  const cuuint64_t strides[3] = {(cuuint64_t)p->nz * 4, (cuuint64_t)comp_stride * 4, (cuuint64_t)comp_stride * 8};
  int tz, rt;
  tma_tile(p, lz, &tz, &rt);
  const cuuint32_t box[4] = {(cuuint32_t)(tz + 4), (cuuint32_t)(rt + 1), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FDTDX_ECUDA, "cuTensorMapEncodeTiled (x halo) failed (" + std::to_string((int)r) + ")");
  p->tmaps[key] = m;
  *out = m;
  return FDTDX_OK;
}

static bool tma_wanted(const FdtdxPlan* p) {
  if (p->use_tma >= 0) return p->use_tma != 0;
  const char* e = getenv("FDTDX_B200_TMA");  // read per launch: tests flip it between plans
  return !(e && e[0] == '0');
}
// The staged kernels cover 128-bit-capable grids whose y / z halos are zero (PML, PEC, PMC faces).
static bool can_tma(const FdtdxPlan* p, const StepParams& P, bool v4) {
  if (!v4 || !tma_wanted(p) || p->wrap[1] || p->wrap[2] || p->bloch) return false;  // Bloch ghosts are mixed at the wrap sites of the marching kernels
  if ((long long)p->nz * p->ny * 4 >= (1LL << 40) || p->ny < 1) return false;
  return encode_tiled_fn() != nullptr;
}
static int tma_chunk(const FdtdxPlan* p, const StepParams& P) {
  int xc = p->xchunk_tma;
  if (xc <= 0) {
    const char* e = getenv("FDTDX_B200_TMA_XCHUNK");
    xc = e ? atoi(e) : 0;
  }
  if (xc <= 0) {
    // Measured on B200 (scripts/sweep_tma.sh): 6-10 planes per CTA is the optimum on 512^3 and on the
    // coupler grid.  Short chunks keep the ~300 resident CTAs inside one narrow x window (same DRAM
    // pages, halo rows shared through L2); the 3-plane ring fill costs less than the drift of long chunks.
    xc = 8;
    // Mid-size grids run only a few waves of the 148 x 2 resident CTAs: pick the chunk length in 5..10
    // whose CTA count wastes the least of its last wave (large grids: no effect).
    int tzc, rtc;
    tma_tile(p, tma_flat_lz(p), &tzc, &rtc);
    const long long tiles = (long long)((p->nz + tzc - 1) / tzc) * ((p->ny + rtc - 1) / rtc);
    const long long nxr = P.x_end - P.x_begin;
    if (tiles * ((nxr + 7) / 8) < 148LL * 2 * 12) {
      // Score = fill of the last wave, discounted when the whole launch is about one wave: then every SM runs its two
      // CTAs in lock step, nothing overlaps their ring fill / drain, and the launch lasts as long as its slowest
      // (CPML-heavy) CTA, whereas several waves of short chunks let the hardware scheduler balance the load.
      // Measured on B200 (scripts/small_grid_sweep.py): C4 grid (135,135,76) 67.4 us/step at 7 planes per CTA,
      // 55.9 at 2; C1 grid 120^3 45.3 at 7 (one full wave), 47.2 at 2.
      double best = -1.0;
      for (int c = 10; c >= 2; --c) {
        const long long ctas = tiles * ((nxr + c - 1) / c);
        const long long waves = (ctas + 148 * 2 - 1) / (148 * 2);
        const double eff = (double)ctas / (double)(waves * 148 * 2) * (1.0 - 0.12 / (double)waves) - 0.004 * std::abs(c - 8);
        if (eff > best) { best = eff; xc = c; }
      }
    }
  }
  return std::max(1, std::min(std::min(xc, 64), P.x_end - P.x_begin));  // 64 = FDTDX_TMA_XC_MAX
}

// Launch with the programmatic-stream-serialization attribute: the kernel may start (up to its
// pdl_wait()) while the previous kernel of the stream drains.  Only for kernels that call pdl_wait()
// before touching anything a previous kernel wrote.  FDTDX_B200_PDL=0 turns the attribute off.
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FDTDX_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// the small per-step kernels (sources, recorder, detectors): opt-in, FDTDX_B200_PDL_AUX=1
static bool pdl_aux_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FDTDX_B200_PDL_AUX");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1 && pdl_enabled();
}
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*k)(KArgs...), dim3 g, dim3 b, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g;
  cfg.blockDim = b;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_aux_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...);  // errors surface through cudaGetLastError()
}

static int launch_sources(FdtdxPlan* p, const StepParams& P, int t, bool is_E, bool rev, cudaStream_t st) {
  if (P.n_src == 0 || P.src_inline) return FDTDX_OK;
  long long mx = 0;
  for (int s = 0; s < P.n_src; ++s) {
    const long long d0 = std::min(P.src_hi[s][0], P.x_end) - std::max(P.src_lo[s][0], P.x_begin);
    if (d0 <= 0) continue;
    mx = std::max(mx, d0 * (P.src_hi[s][1] - P.src_lo[s][1]) * (long long)(P.src_hi[s][2] - P.src_lo[s][2]));
  }
  if (mx <= 0) return FDTDX_OK;
  dim3 g((unsigned)std::min<long long>((mx + 255) / 256, 148 * 8), (unsigned)P.n_src);
  const int te = p->eps_tier == 1 ? 1 : 3, tm = p->mu_tier;
#define GO(A, B) launch_pdl(src_apply_kernel<A, B>, g, dim3(256), st, P, t, is_E ? 1 : 0, rev ? 1 : 0)
  if (te == 1) { if (tm == 0) GO(1, 0); else if (tm == 1) GO(1, 1); else GO(1, 3); }
  else { if (tm == 0) GO(3, 0); else if (tm == 1) GO(3, 1); else GO(3, 3); }
#undef GO
  p->launches++;
  CUDA_TRY(cudaGetLastError());
  return FDTDX_OK;
}

static int launch_E(FdtdxPlan* p, const StepParams& P, int t, bool rev, cudaStream_t st) {
  const bool v4 = can_vec4(p, P);
  const bool sig = p->sigE_tier > 0, ade = p->n_poles > 0, met = p->metric;
  if (rev && ade) return fail(FDTDX_EUNSUPPORTED, "Dispersive time-reversible gradient computation under active development. Use GradientConfig(method='checkpointed') instead.");
  if (rev) {  // update_E_reverse undoes the injection before the update
    int rs = launch_sources(p, P, t, true, true, st);
    if (rs) return rs;
  }
  if (can_tma(p, P, v4)) {
    TmaSet M;
    memset(&M, 0, sizeof(M));
    int rc;
    const int lz = tma_flat_lz(p);
    if ((rc = get_tmap(p, P.H, 3, p->nx, 0, &M.fld_halo, lz))) return rc;
    if ((rc = get_tmap(p, P.E, 3, p->nx, 1, &M.fld_plain, lz))) return rc;
    if ((rc = get_tmap(p, P.eps, p->eps_tier == 1 ? 1 : 3, p->nx, 1, &M.mat_plain, lz))) return rc;
    M.xhalo = M.fld_halo;
    StepParams Q = P;
    Q.xchunk = tma_chunk(p, P);
    Q.flat_lz = lz;
    int tz, rt;
    tma_tile(p, lz, &tz, &rt);
    dim3 g((p->nz + tz - 1) / tz, (p->ny + rt - 1) / rt, (Q.x_end - Q.x_begin + Q.xchunk - 1) / Q.xchunk);
    Q.peer_total = tma_active_warps(p, lz, (int)g.x);  // warps of one x chunk that own rows
    if (lz > 0) CUDA_TRY(fdtdx_dispatch_E4_tmaf(Q, M, t, p->eps_tier, pml_mode(p, P), rev, sig, ade, met, g, st));
    else if (tz == 64) CUDA_TRY(fdtdx_dispatch_E4_tma64(Q, M, t, p->eps_tier, pml_mode(p, P), rev, sig, ade, met, g, st));
    else CUDA_TRY(fdtdx_dispatch_E4_tma(Q, M, t, p->eps_tier, pml_mode(p, P), rev, sig, ade, met, g, st));
  } else {
    // register-marching kernels, four cells per thread: 128-bit accesses (E4) or, on ragged rows, four
    // predicated 32-bit accesses (E1)
    dim3 b(32, p->rows);
    dim3 g((p->nz + 127) / 128, (p->ny + p->rows - 1) / p->rows, (P.x_end - P.x_begin + P.xchunk - 1) / P.xchunk);
    StepParams Q = P;
    Q.peer_total = (int)g.x * p->ny;  // one warp per row
    if (v4) fdtdx_dispatch_E4(Q, t, p->eps_tier, pml_mode(p, P), rev, sig, ade, met, g, b, st);
    else fdtdx_dispatch_E1(Q, t, p->eps_tier, std::min(pml_mode(p, P), 1), rev, sig, ade, met, g, b, st);
  }
  p->launches++;
  CUDA_TRY(cudaGetLastError());
  return rev ? FDTDX_OK : launch_sources(p, P, t, true, false, st);
}

static int launch_H(FdtdxPlan* p, const StepParams& P, int t, bool rev, cudaStream_t st) {
  const bool v4 = can_vec4(p, P);
  if (rev) {
    int rs = launch_sources(p, P, t, false, true, st);
    if (rs) return rs;
  }
  if (can_tma(p, P, v4)) {
    TmaSet M;
    memset(&M, 0, sizeof(M));
    int rc;
    const int lz = tma_flat_lz(p);
    if ((rc = get_tmap(p, P.E, 3, p->nx, 0, &M.fld_halo, lz))) return rc;
    if ((rc = get_tmap(p, P.H, 3, p->nx, 1, &M.fld_plain, lz))) return rc;
    M.mat_plain = M.fld_plain;
    if (p->mu_tier > 0 && (rc = get_tmap(p, P.mu, p->mu_tier == 1 ? 1 : 3, p->nx, 1, &M.mat_plain, lz))) return rc;
    M.xhalo = M.fld_halo;
    if (P.x_hi_mode == 2 && (rc = get_xhalo_tmap(p, P.haloE, P.haloE_cs, &M.xhalo, lz))) return rc;
    StepParams Q = P;
    Q.xchunk = tma_chunk(p, P);
    Q.flat_lz = lz;
    int tz, rt;
    tma_tile(p, lz, &tz, &rt);
    dim3 g((p->nz + tz - 1) / tz, (p->ny + rt - 1) / rt, (Q.x_end - Q.x_begin + Q.xchunk - 1) / Q.xchunk);
    Q.peer_total = tma_active_warps(p, lz, (int)g.x);
    if (lz > 0) CUDA_TRY(fdtdx_dispatch_H4_tmaf(Q, M, t, p->mu_tier, pml_mode(p, P), rev, p->sigH_tier > 0, p->metric, g, st));
    else if (tz == 64) CUDA_TRY(fdtdx_dispatch_H4_tma64(Q, M, t, p->mu_tier, pml_mode(p, P), rev, p->sigH_tier > 0, p->metric, g, st));
    else CUDA_TRY(fdtdx_dispatch_H4_tma(Q, M, t, p->mu_tier, pml_mode(p, P), rev, p->sigH_tier > 0, p->metric, g, st));
  } else {
    dim3 b(32, p->rows);
    dim3 g((p->nz + 127) / 128, (p->ny + p->rows - 1) / p->rows, (P.x_end - P.x_begin + P.xchunk - 1) / P.xchunk);
    StepParams Q = P;
    Q.peer_total = (int)g.x * p->ny;
    if (v4) fdtdx_dispatch_H4(Q, t, p->mu_tier, pml_mode(p, P), rev, p->sigH_tier > 0, p->metric, g, b, st);
    else fdtdx_dispatch_H1(Q, t, p->mu_tier, std::min(pml_mode(p, P), 1), rev, p->sigH_tier > 0, p->metric, g, b, st);
  }
  p->launches++;
  CUDA_TRY(cudaGetLastError());
  return rev ? FDTDX_OK : launch_sources(p, P, t, false, false, st);
}

static void make_grid(const FdtdxPlan* p, GridDev& G) {
  memset(&G, 0, sizeof(G));
  G.nx = p->nx; G.ny = p->ny; G.nz = p->nz; G.x_offset = p->xoff;
  const long long N = (long long)p->nx * p->ny * p->nz;
  for (int a = 0; a < 3; ++a) { G.wrap[a] = p->wrap[a]; G.w[a] = p->d_w[a]; G.sym[a] = p->sym[a]; G.mirror[a] = p->mirror[a]; }
  G.E = (const float*)p->slots[p->e_parity ? FDTDX_SLOT_E_ALT : FDTDX_SLOT_E][0];
  G.H = (const float*)p->slots[p->h_parity ? FDTDX_SLOT_H_ALT : FDTDX_SLOT_H][0];
  G.eps = (const float*)p->slots[FDTDX_SLOT_INV_EPS][0];
  G.mu = (const float*)p->slots[FDTDX_SLOT_INV_MU][0];
  G.eps_cs = (p->eps_tier == 1) ? 0 : N;
  G.mu_cs = (p->mu_tier <= 1) ? 0 : N;
  G.eps_tier = p->eps_tier;
  G.mu_tier = p->mu_tier;
  G.inv_mu_scalar = (float)p->inv_mu_scalar;
  if (p->halo_lo) {
    G.xlo_E = (const float*)p->slots[FDTDX_SLOT_DET_XLO_E][0];
    G.xlo_H = (const float*)p->slots[FDTDX_SLOT_DET_XLO_H][0];
    G.xlo_Hp = (const float*)p->slots[FDTDX_SLOT_DET_XLO_HPREV][0];
    if (!G.xlo_E || !G.xlo_H || !G.xlo_Hp) G.xlo_E = G.xlo_H = G.xlo_Hp = nullptr;
  }
  if (p->bloch) {
    G.Ep = (const float*)p->slots[FDTDX_SLOT_BLOCH_E][0];
    G.Hp = (const float*)p->slots[FDTDX_SLOT_BLOCH_H][0];
    for (int a = 0; a < 3; ++a) { G.bc[a] = p->bloch_c[a]; G.bs[a] = p->bloch_s[a]; }
  }
}

// Upload the detector descriptors (with the currently bound state pointers) for the batched launches.
static int sync_dets(FdtdxPlan* p, cudaStream_t st) {
  if (p->dets.empty()) return FDTDX_OK;
  bool changed = p->dets_dirty;
  {
    // x planes per CTA of det_march_kernel: about one wave of CTAs (148 SMs x 2) over the largest region,
    // 4..DETV_XC_MAX planes each (longer chunks amortise the x-1 rows carried in registers)
    long long tiles = 1, max_ex = 1;
    for (const DetHost& h : p->dets) {
      if (!h.volume) continue;
      tiles = std::max(tiles, (long long)((h.d.hi[2] - h.d.hz0 + DETV_TZ - 1) / DETV_TZ) * ((h.d.hi[1] - h.d.lo[1] + DETV_ROWS - 1) / DETV_ROWS));
      max_ex = std::max<long long>(max_ex, h.d.hi[0] - h.d.lo[0]);
    }
    const long long chunks = std::max(1LL, 296 / tiles);
    int xcl = (int)std::min<long long>(DETV_XC_MAX, std::max(4LL, (max_ex + chunks - 1) / chunks));
    const char* e = getenv("FDTDX_B200_DETV_XC");
    if (e && atoi(e) >= 4 && atoi(e) <= DETV_XC_MAX) xcl = atoi(e);
    p->detv_xcl = xcl;
    for (DetHost& h : p->dets) {
      if (!h.volume || !h.d.part[2]) continue;
      const int want = (h.d.hi[0] - h.d.lo[0] + xcl - 1) / xcl;
      if (h.d.npart[2] != want) { h.d.npart[2] = want; changed = true; }
    }
  }
  for (size_t di = 0; di < p->dets.size(); ++di) {
    DetDev& d = p->dets[di].d;
    float* st[4];
    for (int k = 0; k < 4; ++k) st[k] = (float*)p->slots[FDTDX_SLOT_DET_STATE][4 * di + k];
    for (int k = 0; k < 4; ++k)
      if (st[k] != d.state[k]) { d.state[k] = st[k]; changed = true; }
    if (!d.state[0]) return fail(FDTDX_EUNBOUND, "detector state must be bound");
    if (p->halo_lo && (d.flags & DET_EXACT) && d.lo[0] == 0 &&
        (!p->slots[FDTDX_SLOT_DET_XLO_E][0] || !p->slots[FDTDX_SLOT_DET_XLO_H][0] || !p->slots[FDTDX_SLOT_DET_XLO_HPREV][0]))
      return fail(FDTDX_EUNBOUND, "a detector starts on plane 0 of this x-slab: DET_XLO_E / _H / _HPREV (the lower neighbour's last plane) must be bound");
    {
      const void* eh[3] = {p->slots[FDTDX_SLOT_E][0], p->slots[FDTDX_SLOT_H][0], d.hprev_full};
      bool vol = p->dets[di].volume;
      for (const void* q : eh) vol = vol && (q == nullptr || aligned16(q));
      if (p->e_parity || p->h_parity) vol = vol && aligned16(p->slots[FDTDX_SLOT_E_ALT][0]) && aligned16(p->slots[FDTDX_SLOT_H_ALT][0]);
      const int want = vol ? (d.flags | DET_VOLUME) : (d.flags & ~DET_VOLUME);
      if (want != d.flags) { d.flags = want; changed = true; }
    }
    if (d.kind == FDTDX_DET_ENERGY && (d.flags & DET_SLICES) && (!d.state[1] || !d.state[2]))
      return fail(FDTDX_EUNBOUND, "energy slice detector needs three state buffers");
  }
  if (!changed) return FDTDX_OK;
  std::vector<DetDev>& h = p->h_dets;
  h.clear();
  p->det_max_cells = p->det_max_halo = 0;
  for (auto& d : p->dets) {
    h.push_back(d.d);
    const long long n = (long long)(d.d.hi[0] - d.d.lo[0]) * (d.d.hi[1] - d.d.lo[1]) * (d.d.hi[2] - d.d.lo[2]);
    const long long hn = 3LL * (d.d.hi[0] - d.d.lo[0] + 1) * (d.d.hi[1] - d.d.lo[1] + 1) * (d.d.hi[2] - d.d.lo[2] + 1);
    p->det_max_cells = std::max(p->det_max_cells, n);
    p->det_max_halo = std::max(p->det_max_halo, hn);
  }
  if (!p->d_dets) {
    int rc = to_device(p, h.data(), h.size(), &p->d_dets);
    if (rc) return rc;
  } else {
    CUDA_TRY(cudaMemcpyAsync(p->d_dets, h.data(), h.size() * sizeof(DetDev), cudaMemcpyHostToDevice, st));
  }
  p->dets_dirty = false;
  return FDTDX_OK;
}

static bool any_det_on(const FdtdxPlan* p, int t, bool inverse, bool need_exact, bool* any_post) {
  bool any = false;
  if (any_post) *any_post = false;
  for (const DetHost& h : p->dets) {
    if (((h.d.flags & DET_INVERSE) != 0) != inverse || !h.on[t]) continue;
    if (need_exact && !(h.d.flags & DET_EXACT)) continue;
    any = true;
    if (any_post && ((h.d.flags & DET_REDUCE) || ((h.d.flags & DET_SLICES) && (h.d.flags & DET_SLICE_MEAN)))) *any_post = true;
  }
  return any;
}

static int detectors_gather(FdtdxPlan* p, int t, bool inverse, cudaStream_t st) {
  if (!any_det_on(p, t, inverse, true, nullptr)) return FDTDX_OK;
  int rc = sync_dets(p, st);
  if (rc) return rc;
  GridDev G;
  make_grid(p, G);
  G.xlo_H = G.xlo_Hp;  // this pass copies H BEFORE the H update: the neighbour plane delivered for that instant
  bool any_generic = false, any_volume = false;
  long long vol_rows = 0;
  std::vector<std::array<int, 4>> boxes;
  for (const DetHost& h : p->dets) {
    if (((h.d.flags & DET_INVERSE) != 0) != inverse || !h.on[t] || !(h.d.flags & DET_EXACT)) continue;
    if (h.d.flags & DET_VOLUME) {
      any_volume = true;
      // rows this detector reads: x in [lo-1, hi), y in [lo-1, hi); a wrapped halo row lies on the far face
      std::array<int, 4> box = {std::max(h.d.lo[0] - 1, 0), h.d.hi[0], std::max(h.d.lo[1] - 1, 0), h.d.hi[1]};
      if (h.d.lo[0] == 0 && p->wrap[0]) { box[0] = 0; box[1] = p->nx; }
      if (h.d.lo[1] == 0 && p->wrap[1]) { box[2] = 0; box[3] = p->ny; }
      boxes.push_back(box);
      vol_rows = std::max(vol_rows, 3LL * (h.d.hi[0] - h.d.lo[0] + 1) * (h.d.hi[1] - h.d.lo[1] + 1));
    } else any_generic = true;
  }
  if (any_generic) {
    dim3 g((unsigned)std::min<long long>((p->det_max_halo + 255) / 256, 148 * 8), (unsigned)p->dets.size());
    launch_pdl(det_gather_batch_kernel, g, dim3(256), st, G, (const DetDev*)p->d_dets, t, inverse ? 1 : 0);
    p->launches++;
  }
  p->hprev_fused_t = -1;
  if (any_volume) {
    // forward direction, H half-step on the staged / marching kernels: that kernel stores the H it loads
    // into the plan-wide H_prev scratch (StepParams::hprev_out) - no copy pass.  The time-reversed H
    // kernel un-injects the sources before it loads H, and the full-tensor tier ping-pongs: copy here.
    if (!inverse && !(p->mu_tier == 9 || p->sigH_tier == 9)) {
      p->hprev_fused_t = t;
      if ((int)boxes.size() > FDTDX_MAX_HPBOX) {  // more detectors than box slots: one bounding box
        std::array<int, 4> u = boxes[0];
        for (const auto& b : boxes) { u[0] = std::min(u[0], b[0]); u[1] = std::max(u[1], b[1]); u[2] = std::min(u[2], b[2]); u[3] = std::max(u[3], b[3]); }
        boxes.assign(1, u);
      }
      p->hprev_nbox = (int)boxes.size();
      for (int b = 0; b < p->hprev_nbox; ++b)
        for (int q = 0; q < 4; ++q) p->hprev_box[b][q] = boxes[b][q];
    } else {
      dim3 g((unsigned)std::min<long long>((vol_rows + 7) / 8, 148 * 64), (unsigned)p->dets.size());
      launch_pdl(det_gather_rows_kernel, g, dim3(256), st, G, (const DetDev*)p->d_dets, t, inverse ? 1 : 0);
      p->launches++;
    }
  }
  CUDA_TRY(cudaGetLastError());
  return FDTDX_OK;
}

static int detectors_sample(FdtdxPlan* p, int t, bool inverse, cudaStream_t st) {
  bool any_post = false;
  if (!any_det_on(p, t, inverse, false, &any_post)) return FDTDX_OK;
  int rc = sync_dets(p, st);
  if (rc) return rc;
  GridDev G;
  make_grid(p, G);
  bool any_generic = false, vol[2][2] = {{false, false}, {false, false}};  // [exact][mode]
  int gz = 1, gy = 1, nxc = 1;
  for (const DetHost& h : p->dets) {
    if (((h.d.flags & DET_INVERSE) != 0) != inverse || !h.on[t]) continue;
    if (h.d.flags & DET_VOLUME) {
      vol[(h.d.flags & DET_EXACT) ? 1 : 0][(h.d.kind == FDTDX_DET_ENERGY && G.eps_tier != 9 && G.mu_tier != 9) ? 1 : 0] = true;
      gz = std::max(gz, (h.d.hi[2] - h.d.hz0 + DETV_TZ - 1) / DETV_TZ);
      gy = std::max(gy, (h.d.hi[1] - h.d.lo[1] + DETV_ROWS - 1) / DETV_ROWS);
      nxc = std::max(nxc, (h.d.hi[0] - h.d.lo[0] + p->detv_xcl - 1) / p->detv_xcl);
    } else any_generic = true;
  }
  if (any_generic) {
    dim3 g((unsigned)std::min<long long>((p->det_max_cells + 255) / 256, 148 * 8), (unsigned)p->dets.size());
    launch_pdl(det_sample_batch_kernel, g, dim3(256), st, G, (const DetDev*)p->d_dets, t, inverse ? 1 : 0);
    p->launches++;
  }
  {
    dim3 g(gz, gy, nxc * (unsigned)p->dets.size()), b(32, DETV_ROWS);
    const int inv = inverse ? 1 : 0;
    if (vol[1][1]) {
      if (G.w[0] || G.w[1] || G.w[2]) launch_pdl(det_march_kernel<true, 1, true>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      else launch_pdl(det_march_kernel<true, 1, false>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      p->launches++;
    }
    if (vol[1][0]) {
      if (G.w[0] || G.w[1] || G.w[2]) launch_pdl(det_march_kernel<true, 0, true>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      else launch_pdl(det_march_kernel<true, 0, false>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      p->launches++;
    }
    if (vol[0][1]) {
      if (G.w[0] || G.w[1] || G.w[2]) launch_pdl(det_march_kernel<false, 1, true>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      else launch_pdl(det_march_kernel<false, 1, false>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      p->launches++;
    }
    if (vol[0][0]) {
      if (G.w[0] || G.w[1] || G.w[2]) launch_pdl(det_march_kernel<false, 0, true>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      else launch_pdl(det_march_kernel<false, 0, false>, g, b, st, G, (const DetDev*)p->d_dets, t, inv, nxc, p->detv_xcl);
      p->launches++;
    }
  }
  if (any_post) {
    for (size_t di = 0; di < p->dets.size(); ++di) {
      DetHost& h = p->dets[di];
      if (((h.d.flags & DET_INVERSE) != 0) != inverse || !h.on[t]) continue;
      const long long n = (long long)(h.d.hi[0] - h.d.lo[0]) * (h.d.hi[1] - h.d.lo[1]) * (h.d.hi[2] - h.d.lo[2]);
      if ((h.d.flags & DET_VOLUME) && h.d.kind == FDTDX_DET_ENERGY && (h.d.flags & DET_SLICES) && (h.d.flags & DET_SLICE_MEAN)) {
        // the marching kernel reduced the three means into partial buffers: fold them
        const long long ex = h.d.hi[0] - h.d.lo[0], ey = h.d.hi[1] - h.d.lo[1], ez = h.d.hi[2] - h.d.lo[2];
        const long long nout = ex * ey + ex * ez + ey * ez;
        launch_pdl(det_mean_finish_kernel, dim3((unsigned)((nout + 255) / 256)), dim3(256), st, (const DetDev*)p->d_dets, (int)di, t);
        p->launches++;
        continue;
      }
      if (h.d.flags & DET_REDUCE) {
        const int wpv = (h.d.kind == FDTDX_DET_POYNTING && (h.d.flags & DET_KEEP_ALL)) ? 1 : 0;
        det_reduce_all_kernel<<<h.nvals, 1024, 0, st>>>(h.d, t, h.nvals, wpv);
        p->launches++;
      } else if ((h.d.flags & DET_SLICES) && (h.d.flags & DET_SLICE_MEAN)) {
        const int dims[3] = {h.d.hi[0] - h.d.lo[0], h.d.hi[1] - h.d.lo[1], h.d.hi[2] - h.d.lo[2]};
        for (int axis = 0; axis < 3; ++axis) {
          const long long nout = n / dims[axis];
          det_slice_mean_kernel<<<(unsigned)((nout + 127) / 128), 128, 0, st>>>(h.d, t, axis);
          p->launches++;
        }
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  return FDTDX_OK;
}

static int make_rec(FdtdxPlan* p, RecDev& R) {
  memset(&R, 0, sizeof(R));
  R.dtype = p->rec_dtype;
  R.n_planes = 0;
  for (size_t q = 0; q < p->pmls.size(); ++q) {
    const PmlHost& h = p->pmls[q];
    // x-sharded plans (interfaces/state.py:72-78 shards every recorder array on x): a y / z interface
    // plane is this rank's x-slice of it, an x interface plane lives on the rank that owns that plane
    int cell = (h.dir == 1) ? h.lo_true : h.hi_true - 1;  // boundary.py:134-144
    if (h.axis == 0) {
      cell -= p->xoff;
      if (cell < 0 || cell >= p->nx) continue;
    }
    RecPlane& pl = R.planes[R.n_planes++];
    const int n[3] = {p->nx, p->ny, p->nz - p->zpad};  // the recorder buffers are shaped by the caller's (unpadded) grid
    for (int a = 0; a < 3; ++a) { pl.lo[a] = 0; pl.hi[a] = n[a]; }
    pl.lo[h.axis] = cell; pl.hi[h.axis] = cell + 1;
    pl.data[0] = p->slots[FDTDX_SLOT_REC_DATA][2 * q + 0];
    pl.data[1] = p->slots[FDTDX_SLOT_REC_DATA][2 * q + 1];
    if (!pl.data[0] || !pl.data[1]) return fail(FDTDX_EUNBOUND, "REC_DATA must be bound for every PML slab");
  }
  return FDTDX_OK;
}

#include "tensor_launch.inl"

// ---- peer-memory halo: stream-ordered progress flags ----------------------------------------------
// A half-step that reads a neighbour's boundary plane in place must not start before the neighbour
// has finished the half-step that produced it, and must finish before the neighbour overwrites it.
// Both orders reduce to one rule per half-step: E waits for the low
// neighbour's doneH == (H half-steps issued so far), H waits for the high neighbour's
// doneE == (E half-steps issued so far).  The waits / signals live inside the half-step kernels
// (yee_kernels.cuh: peer_wait_cta / peer_signal_warp): only the CTAs of the chunk that touches the
// shared plane take part, so the interior chunks overlap the wait, the programmatic-dependent-launch
// chain between the half-steps survives, and a whole multi-step run stays one asynchronous
// submission per rank.
// Kernels other than the half-steps that WRITE field planes a neighbour reads in place (interface replay,
// PML field reset of the reverse pass) are bracketed by a stream-ordered fence: wait until both
// neighbours have finished the half-steps issued so far, run the kernel, then advance BOTH progress
// counters (every rank does the same, so the half-step targets stay aligned).
__global__ void peer_fence_wait_kernel(const int* lo_doneH, int targetH, const int* hi_doneE, int targetE, int* err) {
  const int* flag[2] = {lo_doneH, hi_doneE};
  const int target[2] = {targetH, targetE};
  for (int q = 0; q < 2; ++q) {
    if (!flag[q]) continue;
    unsigned ns = 32;
    const long long t0 = clock64();
    for (;;) {
      int v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag[q]) : "memory");
      if (v - target[q] >= 0) break;
      if (clock64() - t0 > (1LL << 36)) { atomicExch(err, 1); break; }
      __nanosleep(ns);
      if (ns < 1024) ns *= 2;
    }
  }
}
__global__ void peer_fence_signal_kernel(int* flags, int valueE, int valueH) {
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flags), "r"(valueE) : "memory");
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flags + 1), "r"(valueH) : "memory");
}
__global__ void peer_flag_set_kernel(int* flag, int value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
static int peer_fence_begin(FdtdxPlan* p, cudaStream_t st) {
  if (!p->peer_mode) return FDTDX_OK;
  peer_fence_wait_kernel<<<1, 1, 0, st>>>(p->halo_lo ? p->peer[0].flags + 1 : nullptr, (int)p->seqH, p->halo_hi ? p->peer[1].flags : nullptr, (int)p->seqE,
                                          p->d_flags + 2);
  CUDA_TRY(cudaGetLastError());
  return FDTDX_OK;
}
static int peer_fence_end(FdtdxPlan* p, cudaStream_t st) {
  if (!p->peer_mode) return FDTDX_OK;
  p->seqE++; p->seqH++;
  peer_fence_signal_kernel<<<1, 1, 0, st>>>(p->d_flags, (int)p->seqE, (int)p->seqH);
  CUDA_TRY(cudaGetLastError());
  return FDTDX_OK;
}

static int step_E(FdtdxPlan* p, int t, int simulate, bool rev, cudaStream_t st) {
  if (p->eps_tier == 9 || p->sigE_tier == 9) return tensor_step(p, t, simulate, rev, /*is_E=*/true, st);
  StepParams P;
  int rc = make_params(p, P, simulate);
  if (rc) return rc;
  // Sources injected by their own O(surface) launch (y / z-normal planes) may touch the plane a neighbour
  // reads; they run outside the half-step kernel, so the in-kernel signal would publish too early (and
  // the reversed pass would un-inject before the wait).  Such plans order whole launches instead.
  const bool whole = p->peer_mode && P.n_src > 0 && !P.src_inline;
  if (p->peer_mode) {
    // in-kernel ordering (common.cuh, StepParams::peer_*): the chunk holding plane 0 waits for the low
    // neighbour's H counter and publishes this rank's E counter; it is scheduled last (z_reverse), a
    // whole kernel after the neighbour's H half-step was issued, so the wait is normally already met
    p->seqE++;
    if (p->halo_lo && !whole) {
      P.peer_wait = p->peer[0].flags + 1;
      P.peer_wait_target = (int)p->seqH;
      P.peer_signal = p->d_flags;
      P.peer_signal_value = (int)p->seqE;
      P.peer_ctr = p->d_flags + 8;
      P.peer_err = p->d_flags + 2;
      P.z_reverse = 1;
    }
    if (whole && p->halo_lo) peer_fence_wait_kernel<<<1, 1, 0, st>>>(p->peer[0].flags + 1, (int)p->seqH, nullptr, 0, p->d_flags + 2);
  }
  rc = launch_E(p, P, t, rev, st);
  if (rc) return rc;
  if (whole) {
    peer_flag_set_kernel<<<1, 1, 0, st>>>(p->d_flags, (int)p->seqE);
    CUDA_TRY(cudaGetLastError());
  }
  if (!rev && p->n_poles > 0) p->p_parity ^= 1;
  return FDTDX_OK;
}

static int step_H(FdtdxPlan* p, int t, int simulate, bool rev, cudaStream_t st) {
  if (p->mu_tier == 9 || p->sigH_tier == 9) return tensor_step(p, t, simulate, rev, /*is_E=*/false, st);
  StepParams P;
  int rc = make_params(p, P, simulate);
  if (rc) return rc;
  const bool whole = p->peer_mode && P.n_src > 0 && !P.src_inline;
  if (p->peer_mode) {
    // the chunk holding plane nx-1 (scheduled last) waits for the high neighbour's E counter and
    // publishes this rank's H counter
    p->seqH++;
    if (p->halo_hi && !whole) {
      P.peer_wait = p->peer[1].flags;
      P.peer_wait_target = (int)p->seqE;
      P.peer_signal = p->d_flags + 1;
      P.peer_signal_value = (int)p->seqH;
      P.peer_ctr = p->d_flags + 9;
      P.peer_err = p->d_flags + 2;
    }
    if (whole && p->halo_hi) peer_fence_wait_kernel<<<1, 1, 0, st>>>(nullptr, 0, p->peer[1].flags, (int)p->seqE, p->d_flags + 2);
  }
  if (!rev && p->hprev_fused_t == t) {
    P.hprev_out = p->d_hprev_full;
    P.hprev_nbox = p->hprev_nbox;
    memcpy(P.hprev_box, p->hprev_box, sizeof(P.hprev_box));
  }
  rc = launch_H(p, P, t, rev, st);
  if (rc) return rc;
  if (whole) {
    peer_flag_set_kernel<<<1, 1, 0, st>>>(p->d_flags + 1, (int)p->seqH);
    CUDA_TRY(cudaGetLastError());
  }
  return FDTDX_OK;
}

// ---- peer-memory halo: CUDA IPC plumbing -----------------------------------------------------------
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetAddressRangeFn address_range_fn() {
  static GetAddressRangeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (GetAddressRangeFn)q;
    cudaGetLastError();
  }
  return fn;
}

extern "C" int fdtdx_b200_peer_export(FdtdxPlan* p, int what, unsigned char* handle64, long long* offset) {
  if (!p || !handle64 || !offset) return fail(FDTDX_EINVAL, "peer_export: null argument");
  void* ptr = nullptr;
  if (what == 0) ptr = p->slots[FDTDX_SLOT_E][0];
  else if (what == 1) ptr = p->slots[FDTDX_SLOT_H][0];
  else if (what == 2) {
    if (!p->d_flags) {
      CUDA_TRY(cudaMalloc((void**)&p->d_flags, 256));
      CUDA_TRY(cudaMemset(p->d_flags, 0, 256));
      p->owned.push_back(p->d_flags);
    }
    ptr = p->d_flags;
  } else return fail(FDTDX_EINVAL, "peer_export: what must be 0 (E), 1 (H) or 2 (flags)");
  if (!ptr) return fail(FDTDX_EUNBOUND, "peer_export: E / H must be bound first");
  GetAddressRangeFn gar = address_range_fn();
  if (!gar) return fail(FDTDX_ECUDA, "cuMemGetAddressRange is not available from this driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (gar(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) return fail(FDTDX_ECUDA, "cuMemGetAddressRange failed");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, (void*)base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = (long long)((CUdeviceptr)ptr - base);
  return FDTDX_OK;
}

static int ipc_open(FdtdxPlan* p, const unsigned char* handle64, void** base) {
  std::string key((const char*)handle64, 64);
  for (auto& kv : p->ipc_open)
    if (kv.first == key) { *base = kv.second; return FDTDX_OK; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* q = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
  p->ipc_open.emplace_back(key, q);
  *base = q;
  return FDTDX_OK;
}

// 0: every neighbour wait of the runs issued so far was met; 1: a wait gave up (results invalid).
// Synchronises the device.
extern "C" int fdtdx_b200_peer_status(FdtdxPlan* p) {
  if (!p) return fail(FDTDX_EINVAL, "peer_status: null plan");
  if (!p->d_flags) return 0;
  int err = 0;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(&err, p->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost));
  return err ? 1 : 0;
}

extern "C" int fdtdx_b200_peer_detach(FdtdxPlan* p) {
  if (!p) return fail(FDTDX_EINVAL, "peer_detach: null plan");
  p->peer_mode = false;
  p->peer[0] = FdtdxPlan::PeerLink();
  p->peer[1] = FdtdxPlan::PeerLink();
  p->tmaps.clear();
  return FDTDX_OK;
}

// side 0: low-x neighbour (its H array + flags), side 1: high-x neighbour (its E array + flags)
extern "C" int fdtdx_b200_peer_attach(FdtdxPlan* p, int side, const unsigned char* field_handle64, long long field_offset,
                                      const unsigned char* flags_handle64, long long flags_offset, int nx_peer) {
  if (!p || side < 0 || side > 1 || !field_handle64 || !flags_handle64 || nx_peer < 1) return fail(FDTDX_EINVAL, "peer_attach: bad argument");
  if ((side == 0 && !p->halo_lo) || (side == 1 && !p->halo_hi)) return fail(FDTDX_EINVAL, "peer_attach: plan has no neighbour on that side (halo_bind)");
  if (!p->d_flags) return fail(FDTDX_EINVAL, "peer_attach: export this rank's flags first");
  void *fb = nullptr, *gb = nullptr;
  int rc;
  if ((rc = ipc_open(p, field_handle64, &fb))) return rc;
  if ((rc = ipc_open(p, flags_handle64, &gb))) return rc;
  p->peer[side].field = (float*)((char*)fb + field_offset);
  p->peer[side].flags = (int*)((char*)gb + flags_offset);
  p->peer[side].nx = nx_peer;
  p->peer_mode = (!p->halo_lo || p->peer[0].field) && (!p->halo_hi || p->peer[1].field);
  p->tmaps.clear();
  return FDTDX_OK;
}

static int step_record(FdtdxPlan* p, int t, int record_detectors, int record_boundaries, cudaStream_t st) {
  if (record_boundaries) {
    if (!p->has_rec) return fail(FDTDX_EINVAL, "Need recorder to record boundaries");
    const int slot = p->slot_of_time[t];
    if (slot >= 0 && !p->pmls.empty()) {
      RecDev R;
      int rc = make_rec(p, R);
      if (rc) return rc;
      GridDev G;
      make_grid(p, G);
      long long fmax = 0;
      for (int q = 0; q < R.n_planes; ++q) {
        const RecPlane& pl = R.planes[q];
        fmax = std::max(fmax, 3LL * (pl.hi[0] - pl.lo[0]) * (pl.hi[1] - pl.lo[1]) * (pl.hi[2] - pl.lo[2]));
      }
      if (R.n_planes > 0) {
        dim3 g((unsigned)std::min<long long>((fmax + 255) / 256, 1024), 2 * R.n_planes);
        launch_pdl(rec_record_kernel, g, dim3(256), st, R, (float*)G.E, (float*)G.H, p->nx, p->ny, p->nz, slot);
        p->launches++;
        CUDA_TRY(cudaGetLastError());
      }
    }
  }
  if (record_detectors) return detectors_sample(p, t, false, st);
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_run_forward_phase(FdtdxPlan* p, int t, int phase, int record_detectors,
                                            int record_boundaries, int simulate_boundaries, void* stream) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  if (t < 0 || t >= p->T) return fail(FDTDX_EINVAL, "time step outside [0, T)");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = finalize(p);
  if (rc) return rc;
  if (phase == 0) {
    if (record_detectors && (rc = detectors_gather(p, t, false, st))) return rc;
    return step_E(p, t, simulate_boundaries, false, st);
  }
  if (phase == 1) return step_H(p, t, simulate_boundaries, false, st);
  if (phase == 3) return record_detectors ? detectors_gather(p, t, false, st) : FDTDX_OK;
  return step_record(p, t, record_detectors, record_boundaries, st);
}

extern "C" int fdtdx_b200_run_half_range(FdtdxPlan* p, int t, int which, int x_begin, int x_end, int simulate_boundaries,
                                         void* stream) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  if (t < 0 || t >= p->T) return fail(FDTDX_EINVAL, "time step outside [0, T)");
  if (x_begin < 0 || x_end > p->nx || x_begin > x_end) return fail(FDTDX_EINVAL, "run_half_range: bad plane range");
  if (x_begin == x_end) return FDTDX_OK;
  if (p->eps_tier == 9 || p->mu_tier == 9 || p->sigE_tier == 9 || p->sigH_tier == 9)
    return fail(FDTDX_EUNSUPPORTED, "run_half_range: full-tensor media are not supported on x-sharded plans");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = finalize(p);
  if (rc) return rc;
  StepParams P;
  if ((rc = make_params(p, P, simulate_boundaries))) return rc;
  P.x_begin = x_begin;
  P.x_end = x_end;
  if (which == 0) {
    rc = launch_E(p, P, t, false, st);
    // ADE ping-pong flips once per step: callers issue the range starting at plane 0 last
    if (!rc && p->n_poles > 0 && x_begin == 0) p->p_parity ^= 1;
    return rc;
  }
  if (p->hprev_fused_t == t) {
    P.hprev_out = p->d_hprev_full;
    P.hprev_nbox = p->hprev_nbox;
    memcpy(P.hprev_box, p->hprev_box, sizeof(P.hprev_box));
  }
  return launch_H(p, P, t, false, st);
}

extern "C" int fdtdx_b200_total_energy(FdtdxPlan* p, float* d_out, void* stream) {
  if (!p || !d_out) return fail(FDTDX_EINVAL, "total_energy: null argument");
  if (p->eps_tier == 9 || p->mu_tier == 9) return fail(FDTDX_EUNSUPPORTED, "total_energy: full-tensor media are not supported");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = finalize(p);
  if (rc) return rc;
  GridDev G;
  make_grid(p, G);
  if (!G.E || !G.H || !G.eps) return fail(FDTDX_EUNBOUND, "E, H and INV_EPS must be bound");
  if (!p->d_energy_partial && (rc = to_device<double>(p, nullptr, FDTDX_ENERGY_BLOCKS, &p->d_energy_partial))) return rc;
  energy_partial_kernel<<<FDTDX_ENERGY_BLOCKS, 256, 0, st>>>(G, p->d_energy_partial, p->zpad);
  energy_final_kernel<<<1, 256, 0, st>>>(p->d_energy_partial, FDTDX_ENERGY_BLOCKS, d_out);
  p->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_get_xchunk(FdtdxPlan* p) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  StepParams P;
  int rc = finalize(p);
  if (rc) return rc;
  if ((rc = make_params(p, P, 1))) return rc;
  return P.xchunk;
}

extern "C" int fdtdx_b200_run_forward(FdtdxPlan* p, int t0, int n, int record_detectors, int record_boundaries,
                                      int simulate_boundaries, void* stream) {
  for (int t = t0; t < t0 + n; ++t)
    for (int phase = 0; phase < 3; ++phase) {
      int rc = fdtdx_b200_run_forward_phase(p, t, phase, record_detectors, record_boundaries, simulate_boundaries, stream);
      if (rc) return rc;
    }
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_run_reverse_phase(FdtdxPlan* p, int t, int phase, int record_detectors, int reset_fields, void* stream) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = finalize(p);
  if (rc) return rc;
  if (t < 0 || t >= p->T) return fail(FDTDX_EINVAL, "reverse time step outside [0, T)");
  if (!p->has_rec) return fail(FDTDX_EINVAL, "Need recorder to record boundaries");
  if ((p->halo_lo || p->halo_hi) && !p->peer_mode)
    return fail(FDTDX_EUNSUPPORTED, "run_reverse on an x-sharded plan needs the peer-memory halo (peer_attach)");
  if (phase == 0) {  // add_interfaces (update.py:1181-1222)
    if (p->pmls.empty()) return FDTDX_OK;
    RecDev R;
    if ((rc = make_rec(p, R))) return rc;
    GridDev G;
    make_grid(p, G);
    long long fmax = 0;
    for (int q = 0; q < R.n_planes; ++q) {
      const RecPlane& pl = R.planes[q];
      fmax = std::max(fmax, 3LL * (pl.hi[0] - pl.lo[0]) * (pl.hi[1] - pl.lo[1]) * (pl.hi[2] - pl.lo[2]));
    }
    if ((rc = peer_fence_begin(p, st))) return rc;
    if (R.n_planes > 0) {
      dim3 g((unsigned)std::min<long long>((fmax + 255) / 256, 1024), 2 * R.n_planes);
      launch_pdl(rec_replay_kernel, g, dim3(256), st, R, (float*)G.E, (float*)G.H, p->nx, p->ny, p->nz, (int)p->replay_a[t],
                 (int)p->replay_b[t], p->replay_w[t]);
      p->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    return peer_fence_end(p, st);
  }
  if (phase == 1) return record_detectors ? detectors_gather(p, t, true, st) : FDTDX_OK;
  if (phase == 2) return step_H(p, t, 0, true, st);
  if (phase == 3) return step_E(p, t, 0, true, st);
  if (phase == 4) {  // apply_field_reset in every PML slab (backward.py:117-122)
    if (!reset_fields || p->pmls.empty()) return FDTDX_OK;
    BoxList B;
    B.n = 0;
    long long nmax = 0;
    const int nn[3] = {p->nx, p->ny, p->nz};
    for (const PmlHost& h : p->pmls) {
      for (int a = 0; a < 3; ++a) { B.lo[B.n][a] = 0; B.hi[B.n][a] = nn[a]; }
      int lo = h.lo_true, hi = h.hi_true;  // the reference's slab, not a padded superset of it
      if (h.axis == 0) {  // this rank's part of an x slab
        lo = std::max(lo - p->xoff, 0); hi = std::min(hi - p->xoff, p->nx);
        if (hi <= lo) continue;
      }
      B.lo[B.n][h.axis] = lo; B.hi[B.n][h.axis] = hi;
      nmax = std::max(nmax, (long long)(hi - lo) * nn[(h.axis + 1) % 3] * nn[(h.axis + 2) % 3]);
      B.n++;
    }
    GridDev G;
    make_grid(p, G);
    if ((rc = peer_fence_begin(p, st))) return rc;
    if (B.n > 0) {
      dim3 g((unsigned)std::min<long long>((nmax + 255) / 256, 4096), B.n);
      launch_pdl(reset_pml_kernel, g, dim3(256), st, B, (float*)G.E, (float*)G.H, p->nx, p->ny, p->nz);
      p->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    return peer_fence_end(p, st);
  }
  if (phase == 5) return record_detectors ? detectors_sample(p, t, true, st) : FDTDX_OK;
  return fail(FDTDX_EINVAL, "run_reverse_phase: phase must be 0..5");
}

extern "C" int fdtdx_b200_run_reverse(FdtdxPlan* p, int t_from, int n, int record_detectors, int reset_fields,
                                      void* stream) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  for (int t = t_from - 1; t > t_from - 1 - n; --t)
    for (int phase = 0; phase < 6; ++phase) {
      int rc = fdtdx_b200_run_reverse_phase(p, t, phase, record_detectors, reset_fields, stream);
      if (rc) return rc;
    }
  return FDTDX_OK;
}

static bool adj_fused_wanted() {
  const char* e = getenv("FDTDX_B200_ADJ_FUSED");  // 0: the two-kernel form (adj_local4 + adj_gather4) everywhere
  return !(e && e[0] == '0');
}
static int adj_fused_xchunk() {
  const char* e = getenv("FDTDX_B200_ADJ_XC");
  return e ? atoi(e) : 0;
}
static bool adj_det_streams_wanted() {
  const char* e = getenv("FDTDX_B200_ADJ_DET_STREAMS");  // 0: the detectors' cotangent kernels run one after the other
  return !(e && e[0] == '0');
}
static bool adj_async_wanted() {
  const char* e = getenv("FDTDX_B200_ADJ_ASYNC");  // 0: the fused kernel loads each plane directly instead of staging the next one with cp.async
  return !(e && e[0] == '0');
}
static bool adj_flat_wanted() {
  const char* e = getenv("FDTDX_B200_ADJ_FLAT");  // 0: a warp per row also on thin grids
  return !(e && e[0] == '0');
}
static bool adj_interleave_wanted() {
  const char* e = getenv("FDTDX_B200_ADJ_INTERLEAVE");  // 0: reverse step, then recompute the step into scratch
  return !(e && e[0] == '0');
}
static size_t pml_slab_cells(const FdtdxPlan* p, const PmlHost& h) {
  int lo = h.lo, hi = h.hi;
  if (h.axis == 0) { lo = std::max(lo - p->xoff, 0); hi = std::min(hi - p->xoff, p->nx); }
  const long long len = std::max(hi - lo, 0);
  const long long nn[3] = {p->nx, p->ny, p->nz};
  return (size_t)(len * nn[(h.axis + 1) % 3] * nn[(h.axis + 2) % 3]);
}
// the fused adjoint kernel leaves the psi cotangents in the plan-owned partner buffers after an odd number of
// half-steps: bring them home (the caller reads the bound buffers)
static int adj_cotpsi_home(FdtdxPlan* p, cudaStream_t st) {
  for (int kind = 0; kind < 2; ++kind) {
    if (!p->cotpsi_parity[kind]) continue;
    const int cot_slot = kind ? FDTDX_SLOT_COT_PSI_E : FDTDX_SLOT_COT_PSI_H;
    for (size_t q = 0; q < p->pmls.size(); ++q)
      for (int w = 0; w < 2; ++w) {
        float* bound = (float*)p->slots[cot_slot][2 * q + w];
        if (bound && p->d_cotpsi_alt[kind].size() > 2 * q + w)
          CUDA_TRY(cudaMemcpyAsync(bound, p->d_cotpsi_alt[kind][2 * q + w], pml_slab_cells(p, p->pmls[q]) * 4, cudaMemcpyDeviceToDevice, st));
      }
    p->cotpsi_parity[kind] = 0;
  }
  return FDTDX_OK;
}

static int adjoint_half(FdtdxPlan* p, const StepParams& S, bool is_E, const float* F, const float* G, float* lamF, float* lamG,
                        const float* lam_extra, cudaStream_t st) {
  AdjParams A;
  memset(&A, 0, sizeof(A));
  const long long N = (long long)p->nx * p->ny * p->nz;
  A.nx = p->nx; A.ny = p->ny; A.nz = p->nz;
  for (int a = 0; a < 3; ++a) { A.wrap[a] = p->wrap[a]; A.pml[a] = S.pml[a]; A.sc[a] = is_E ? p->d_sB[a] : p->d_sF[a]; }
  A.cour = S.cour; A.eta0 = S.eta0; A.mat_scalar = S.inv_mu_scalar; A.is_E = is_E ? 1 : 0;
  A.F = F; A.G = G; A.lamF = lamF; A.lamG = lamG; A.lam_extra = lam_extra; A.ld = p->d_ld;
  if (is_E) {
    A.mat = S.eps; A.mat_tier = p->eps_tier; A.mat_cs = (p->eps_tier == 1) ? 0 : N;
    A.sig = S.sigE; A.sig_cs = S.sigE_cs;
    A.g_mat = (float*)p->slots[FDTDX_SLOT_GRAD_INV_EPS][0];
  } else {
    A.mat = S.mu; A.mat_tier = p->mu_tier; A.mat_cs = (p->mu_tier <= 1) ? 0 : N;
    A.sig = S.sigH; A.sig_cs = S.sigH_cs;
    A.g_mat = (p->mu_tier > 0) ? (float*)p->slots[FDTDX_SLOT_GRAD_INV_MU][0] : nullptr;
  }
  const int kind = is_E ? 1 : 0;
  const int cot_slot = is_E ? FDTDX_SLOT_COT_PSI_E : FDTDX_SLOT_COT_PSI_H;
  for (size_t q = 0; q < p->pmls.size(); ++q) {
    const PmlHost& h = p->pmls[q];
    for (int w = 0; w < 2; ++w) {
      float* bound = (float*)p->slots[cot_slot][2 * q + w];
      const bool alt = bound && p->cotpsi_parity[kind] && p->d_cotpsi_alt[kind].size() > 2 * q + w;
      A.lam_psi[h.axis][h.dir][w] = alt ? p->d_cotpsi_alt[kind][2 * q + w] : bound;
    }
  }
  A.n_walls = S.n_walls; A.walls = S.walls;
  if (is_E && p->n_poles > 0) {
    A.n_poles = p->n_poles; A.has_c4 = p->has_c4; A.c_cs = S.c_cs;
    A.Pc = S.P_cur; A.Pq = S.P_new;  // the input state's P_curr / P_prev (parity as bound)
    A.cf[0] = S.c1; A.cf[1] = S.c2; A.cf[2] = S.c3; A.cf[3] = S.c4;
    A.lamP = (float*)p->slots[FDTDX_SLOT_COT_P][0];
    A.lamQ = (float*)p->slots[FDTDX_SLOT_COT_P_PREV][0];
    if (!A.lamP || !A.lamQ) return fail(FDTDX_EUNBOUND, "COT_P / COT_P_PREV must be bound for the adjoint of a dispersive step");
    for (int k = 0; k < 4; ++k) A.g_c[k] = (float*)p->slots[FDTDX_SLOT_GRAD_C1 + k][0];
  }
  // 128-bit form when every (., N)-shaped operand is 16-byte aligned and rows are multiples of 4 cells
  bool v4 = (p->nz % 4 == 0);
  const void* ptrs[] = {F, G, lamF, lamG, lam_extra, A.ld, A.mat, A.sig, A.g_mat, A.sc[2]};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) v4 = false;
  // Single-pass form (adj_fused4_kernel): whenever lambda' is left as it is by the half-step's transpose, or only
  // masked by PEC / PMC walls - no conductivity, no ADE, no extra input cotangent.
  int walls_kind = 0;
  for (const WallDev& wl : p->walls) walls_kind += (wl.kind == (is_E ? 0 : 1)) ? 1 : 0;
  // the fused kernel takes one z-slab side per 4-cell group: keep degenerate grids (both z slabs in one group) off it
  const AxisPmlDev& zs = A.pml[2];
  const bool z_mixed = zs.lo_len > 0 && zs.hi_start < p->nz && zs.hi_start <= ((zs.lo_len - 1) / 4) * 4 + 3;
  const bool fused = v4 && adj_fused_wanted() && !A.sig && !lam_extra && A.n_poles == 0 && walls_kind <= FDTDX_ADJ_MAXW && !z_mixed;
  if (fused) {
    // ping-pong partner of every bound psi-cotangent buffer (a neighbour's re-evaluation reads the old value)
    std::vector<float*>& alt = p->d_cotpsi_alt[kind];
    if (alt.size() < 2 * p->pmls.size()) {
      alt.assign(2 * p->pmls.size(), nullptr);
      for (size_t q = 0; q < p->pmls.size(); ++q) {
        int rcq;
        for (int w = 0; w < 2; ++w)
          if ((rcq = to_device<float>(p, nullptr, pml_slab_cells(p, p->pmls[q]), &alt[2 * q + w]))) return rcq;
      }
    }
    for (size_t q = 0; q < p->pmls.size(); ++q) {
      const PmlHost& h = p->pmls[q];
      for (int w = 0; w < 2; ++w) {
        float* bound = (float*)p->slots[cot_slot][2 * q + w];
        A.lam_psi_new[h.axis][h.dir][w] = !bound ? nullptr : (p->cotpsi_parity[kind] ? bound : alt[2 * q + w]);
      }
    }
    // thin rows (Nz <= 124): the CTA's threads laid over (row, z quad) pairs, 256 / lz rows per CTA
    const int flz = (adj_flat_wanted() && p->nz / 4 >= 5 && p->nz / 4 <= 31) ? p->nz / 4 : 0;
    const int rt = flz ? 256 / flz : 8;
    A.flat_lz = flz;
    const long long tiles = (long long)((p->nz + 127) / 128) * ((p->ny + rt - 1) / rt);
    int xc = adj_fused_xchunk();
    if (xc <= 0) xc = (int)std::max<long long>(4, std::min<long long>(16, (long long)p->nx * tiles / (148 * 2 * 3)));
    A.xchunk = std::min(xc, p->nx);
    const bool grad = A.g_mat != nullptr;
    const bool met = A.sc[0] || A.sc[1] || A.sc[2];
    dim3 b(32, 8), g((p->nz + 127) / 128, (p->ny + rt - 1) / rt, (p->nx + A.xchunk - 1) / A.xchunk);
    A.psi_vec = 1;
    for (int a = 0; a < 2; ++a)
      for (int sd = 0; sd < 2; ++sd)
        for (int w = 0; w < 2; ++w) {
          const void* ps[] = {A.lam_psi[a][sd][w], A.lam_psi_new[a][sd][w], is_E ? A.pml[a].psiE[sd][w] : A.pml[a].psiH[sd][w]};
          for (const void* q : ps)
            if (q && !aligned16(q)) A.psi_vec = 0;
        }
    // one pass for the cotangents, then (when a material gradient is wanted) one for g += lambda_in . (+-c K)
    const bool use_async = adj_async_wanted();
#define GO_F(IE, MT, ME) do { \
      if (use_async) { \
        auto k = adj_fused4_kernel<IE, MT, ME, 8, true>; \
        constexpr int smem = 2 * adj_stage_vecs<MT>() * 256 * 16; \
        static bool attr_set = false; \
        if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; } \
        k<<<g, b, smem, st>>>(A); \
      } else { \
        adj_fused4_kernel<IE, MT, ME, 8, false><<<g, b, 0, st>>>(A); \
      } \
    } while (0)
#define GO(IE, MT) do { \
      if (met) GO_F(IE, MT, true); else GO_F(IE, MT, false); \
      if (grad) { if (met) adj_grad4_kernel<IE, MT, true, 8><<<g, b, 0, st>>>(A); else adj_grad4_kernel<IE, MT, false, 8><<<g, b, 0, st>>>(A); } \
    } while (0)
    if (is_E) { if (A.mat_tier == 3) GO(true, 3); else GO(true, 1); }
    else { if (A.mat_tier == 0) GO(false, 0); else if (A.mat_tier == 1) GO(false, 1); else GO(false, 3); }
#undef GO
#undef GO_F
    p->launches += grad ? 1 : 0;
    p->cotpsi_parity[kind] ^= 1;
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return FDTDX_OK;
  }
  // two-kernel form: the six derivative cotangents go through a 6 N scratch (allocated on first use only - the fused
  // kernels never touch it)
  if (!p->d_ld) {
    int rcl = to_device<float>(p, nullptr, (size_t)6 * N, &p->d_ld);
    if (rcl) return rcl;
    if (!aligned16(p->d_ld)) v4 = false;
  }
  A.ld = p->d_ld;
  if (v4) {
    dim3 b(32, 8), g((p->nz + 127) / 128, (p->ny + 7) / 8, p->nx);
    if (is_E) { adj_local4_kernel<true><<<g, b, 0, st>>>(A); adj_gather4_kernel<true><<<g, b, 0, st>>>(A); }
    else { adj_local4_kernel<false><<<g, b, 0, st>>>(A); adj_gather4_kernel<false><<<g, b, 0, st>>>(A); }
  } else {
    const unsigned blocks = (unsigned)((N + 255) / 256);
    adj_local_kernel<<<blocks, 256, 0, st>>>(A);
    adj_gather_kernel<<<blocks, 256, 0, st>>>(A);
  }
  p->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return FDTDX_OK;
}

extern "C" int fdtdx_b200_run_adjoint(FdtdxPlan* p, int t_from, int n, void* stream) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = finalize(p);
  if (rc) return rc;
  if (p->nx != p->nxg) return fail(FDTDX_EUNSUPPORTED, "run_adjoint on x-sharded plans is not supported yet");
  if (p->eps_tier == 9 || p->mu_tier == 9 || p->sigE_tier == 9 || p->sigH_tier == 9)
    return fail(FDTDX_EUNSUPPORTED, "run_adjoint: full-tensor media are not supported yet");
  if (p->n_poles > 0 && !p->adjoint_exact)
    return fail(FDTDX_EUNSUPPORTED, "Dispersive time-reversible gradient computation under active development. Use GradientConfig(method='checkpointed') instead.");
  const long long N = (long long)p->nx * p->ny * p->nz;
  float* lamE = (float*)p->slots[FDTDX_SLOT_COT_E][0];
  float* lamH = (float*)p->slots[FDTDX_SLOT_COT_H][0];
  if (!lamE || !lamH || !p->slots[FDTDX_SLOT_GRAD_INV_EPS][0]) return fail(FDTDX_EUNBOUND, "COT_E, COT_H and GRAD_INV_EPS must be bound");
  if (!p->d_Htmp) {  // scratch of the interleaved iteration; the re-run order adds a copy of E on first use
    if ((rc = to_device<float>(p, nullptr, (size_t)3 * N, &p->d_Htmp))) return rc;
    if ((rc = to_device<float>(p, nullptr, (size_t)3 * N, &p->d_lamHx))) return rc;
  }
  // Reversible mode, interleaved (default): the reversed step is split around the two transposes, so every
  // state a transpose needs is in the field arrays at the moment it runs and nothing is recomputed -
  //   replay interfaces -> [keep H_{t+1} for the detector cotangents] -> reverse H (H := H_t) ->
  //   detector + H half-step transposes (E_{t+1}, H_t) -> reverse E (E := E_t) -> E half-step transpose (E_t, H_t).
  // E_{t+1} / H_{t+1} are then the reconstructed values instead of a re-run of the forward step from (E_t, H_t)
  // with the frozen psi: identical outside the CPML slabs up to rounding, and inside them the reference's
  // reversible gradient is reconstruction noise either way (tests compare outside the slabs).
  // FDTDX_B200_ADJ_INTERLEAVE=0 and the exact mode use the reverse-then-recompute order.
  const bool interleave = !p->adjoint_exact && adj_interleave_wanted();
  // does the H half-step transpose run fused (then the detector H_prev cotangent is added afterwards, box by box)?
  int pmc_walls = 0;
  for (const WallDev& wl : p->walls) pmc_walls += (wl.kind == 1) ? 1 : 0;
  const AxisPmlDev& zsl = p->axis[2];
  const bool z_mixed = zsl.lo_len > 0 && zsl.hi_start < p->nz && zsl.hi_start <= ((zsl.lo_len - 1) / 4) * 4 + 3;
  const bool fusedH = adj_fused_wanted() && (p->nz % 4 == 0) && p->sigH_tier == 0 && pmc_walls <= FDTDX_ADJ_MAXW && !z_mixed;
  bool fusedH_checked = false;
  for (int t = t_from - 1; t > t_from - 1 - n; --t) {
    if (t < 0) break;  // the reference's extra t = -1 iteration (fdtd.py:253-260) starts from the zero state; skipped
    bool any_on = false;
    for (size_t di = 0; di < p->dets.size(); ++di) {
      DetHost& h = p->dets[di];
      if ((h.d.flags & DET_INVERSE) || !h.on[t]) continue;
      for (int k = 0; k < 4; ++k) any_on = any_on || p->slots[FDTDX_SLOT_COT_DET][4 * di + k] != nullptr;
    }
    // Energy / Poynting cotangents are linearised at E_{t+1}, H_{t+1}: where such a detector reaches into a CPML
    // slab the reference's re-run of the step (frozen psi) and the reconstructed fields differ by more than
    // rounding, so the steps on which one is on (with a cotangent) keep the reverse-then-recompute order.
    bool interleave_t = interleave;
    for (size_t di = 0; interleave_t && di < p->dets.size(); ++di) {
      DetHost& h = p->dets[di];
      if ((h.d.flags & DET_INVERSE) || !h.on[t] || !(h.d.kind == 1 || h.d.kind == 2)) continue;
      bool any_cot = false;
      for (int k = 0; k < 4; ++k) any_cot = any_cot || p->slots[FDTDX_SLOT_COT_DET][4 * di + k] != nullptr;
      if (!any_cot) continue;
      for (const PmlHost& m : p->pmls) {
        const int lo = m.lo - (m.axis == 0 ? p->xoff : 0), hi = m.hi - (m.axis == 0 ? p->xoff : 0);
        if (h.d.lo[m.axis] - 1 < hi && h.d.hi[m.axis] + 1 > lo) interleave_t = false;
      }
    }
    StepParams S;
    const float* E1;  // E_{t+1}
    const float* H1;  // H_{t+1}
    if (interleave_t) {
      if ((rc = fdtdx_b200_run_reverse_phase(p, t, 0, 0, 0, stream))) return rc;
      if ((rc = make_params(p, S, 1))) return rc;
      if (any_on) CUDA_TRY(cudaMemcpyAsync(p->d_Htmp, S.H, (size_t)3 * N * 4, cudaMemcpyDeviceToDevice, st));
      if ((rc = fdtdx_b200_run_reverse_phase(p, t, 2, 0, 0, stream))) return rc;
      E1 = S.E;
      H1 = p->d_Htmp;
    } else {
      if (!p->d_Etmp && (rc = to_device<float>(p, nullptr, (size_t)3 * N, &p->d_Etmp))) return rc;
      // (1) reconstruct the state at t (backward.py:62-135, record_detectors=False, reset_fields=False);
      //     exact mode: the caller has bound the stored state of step t instead
      if (!p->adjoint_exact && (rc = fdtdx_b200_run_reverse(p, t + 1, 1, 0, 0, stream))) return rc;
      if ((rc = make_params(p, S, 1))) return rc;
      // (2) recompute E_{t+1}, H_{t+1} from the reconstructed state without touching psi.  The staged
      // kernels write straight into the scratch buffers; the marching kernels update in place, so their
      // input is copied first.
      StepParams F1 = S;
      F1.psi_store = 0;
      F1.p_store = 0;
      StepParams F2;
      if (can_tma(p, S, can_vec4(p, S))) {
        F1.E_out = p->d_Etmp;
        if ((rc = launch_E(p, F1, t, false, st))) return rc;
        F2 = F1;
        F2.E = p->d_Etmp;
        F2.E_out = nullptr;
        F2.H_out = p->d_Htmp;
        if ((rc = launch_H(p, F2, t, false, st))) return rc;
      } else {
        CUDA_TRY(cudaMemcpyAsync(p->d_Etmp, S.E, (size_t)3 * N * 4, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(p->d_Htmp, S.H, (size_t)3 * N * 4, cudaMemcpyDeviceToDevice, st));
        F1.E = p->d_Etmp;
        if ((rc = launch_E(p, F1, t, false, st))) return rc;
        F2 = F1;
        F2.H = p->d_Htmp;
        if ((rc = launch_H(p, F2, t, false, st))) return rc;
      }
      E1 = p->d_Etmp;
      H1 = p->d_Htmp;
    }
    float* E = S.E;
    float* H = S.H;  // H_t
    // (3) detector cotangents at step t
    bool any_det = false;
    std::vector<std::array<int, 6>> boxes;  // per detector: the box its H_prev cotangents were scattered into
    const int nn[3] = {p->nx, p->ny, p->nz};
    std::vector<size_t> active;
    for (size_t di = 0; di < p->dets.size(); ++di) {
      DetHost& h = p->dets[di];
      if ((h.d.flags & DET_INVERSE) || !h.on[t]) continue;
      // a detector is skipped only when none of its state leaves carries a cotangent
      bool any_cot = false;
      for (int k = 0; k < 4; ++k) any_cot = any_cot || p->slots[FDTDX_SLOT_COT_DET][4 * di + k] != nullptr;
      if (any_cot) active.push_back(di);
    }
    // two-kernel form: the scratch is consumed whole, so it is zeroed whole; fused form: it is kept zero
    // outside the boxes by adj_box_add_clear_kernel
    if (!active.empty() && !fusedH) CUDA_TRY(cudaMemsetAsync(p->d_lamHx, 0, (size_t)3 * N * 4, st));
    const bool fork = active.size() > 1 && adj_det_streams_wanted();
    if (fork) {
      if (!p->adj_fork) {
        CUDA_TRY(cudaEventCreateWithFlags(&p->adj_fork, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
          CUDA_TRY(cudaStreamCreateWithFlags(&p->adj_side[k], cudaStreamNonBlocking));
          CUDA_TRY(cudaEventCreateWithFlags(&p->adj_join[k], cudaEventDisableTiming));
        }
      }
      CUDA_TRY(cudaEventRecord(p->adj_fork, st));
      for (int k = 0; k < 2; ++k) CUDA_TRY(cudaStreamWaitEvent(p->adj_side[k], p->adj_fork, 0));
    }
    for (size_t ai = 0; ai < active.size(); ++ai) {
      const size_t di = active[ai];
      DetHost& h = p->dets[di];
      cudaStream_t ds = fork ? p->adj_side[ai % 2] : st;
      any_det = true;
      std::array<int, 6> bx;
      for (int a = 0; a < 3; ++a) {  // the stencil reaches one cell down in x, y and one up in z (+ wrap)
        int lo = h.d.lo[a] - 1, hi = h.d.hi[a] + 1;
        if (p->wrap[a] && (lo < 0 || hi > nn[a])) { lo = 0; hi = nn[a]; }
        bx[2 * a] = std::max(lo, 0);
        bx[2 * a + 1] = std::min(hi, nn[a]);
      }
      boxes.push_back(bx);
      GridDev G;
      make_grid(p, G);
      G.E = E1;
      G.H = H;  // H_prev gather reads the step's input H
      if (h.d.flags & DET_EXACT) {
        const long long hn = 3LL * (h.d.hi[0] - h.d.lo[0] + 1) * (h.d.hi[1] - h.d.lo[1] + 1) * (h.d.hi[2] - h.d.lo[2] + 1);
        det_gather_hprev_kernel<<<(int)std::min<long long>((hn + 255) / 256, 148 * 16), 256, 0, ds>>>(G, h.d);
        p->launches++;
      }
      G.H = H1;
      DetAdj A;
      for (int k = 0; k < 4; ++k) A.cot[k] = (const float*)p->slots[FDTDX_SLOT_COT_DET][4 * di + k];
      A.lamE = lamE; A.lamH = lamH; A.lamHprev = p->d_lamHx;
      A.g_eps = (float*)p->slots[FDTDX_SLOT_GRAD_INV_EPS][0];
      A.eps_tier = p->eps_tier;
      A.g_mu = (p->mu_tier > 0) ? (float*)p->slots[FDTDX_SLOT_GRAD_INV_MU][0] : nullptr;
      A.mu_tier = p->mu_tier;
      const long long dn = (long long)(h.d.hi[0] - h.d.lo[0]) * (h.d.hi[1] - h.d.lo[1]) * (h.d.hi[2] - h.d.lo[2]);
      det_adjoint_kernel<<<(unsigned)((dn + 127) / 128), 128, 0, ds>>>(G, h.d, A, t);
      p->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    if (fork)
      for (int k = 0; k < 2; ++k) {
        CUDA_TRY(cudaEventRecord(p->adj_join[k], p->adj_side[k]));
        CUDA_TRY(cudaStreamWaitEvent(st, p->adj_join[k], 0));
      }
    // (4) H half-step transpose: lambda_H' -> lambda_H_in, accumulates into lambda_E'
    if (fusedH && !fusedH_checked) {  // same alignment test as adjoint_half's (only matters before any scatter happened)
      const void* ptrs[] = {H, E1, lamH, lamE, S.mu, p->mu_tier > 0 ? p->slots[FDTDX_SLOT_GRAD_INV_MU][0] : nullptr, p->d_sF[2]};
      for (const void* q : ptrs)
        if (q && !aligned16(q)) return fail(FDTDX_EINVAL, "run_adjoint: unaligned operands with a multiple-of-4 Nz (set FDTDX_B200_ADJ_FUSED=0)");
      fusedH_checked = true;
    }
    if ((rc = adjoint_half(p, S, false, H, E1, lamH, lamE, (any_det && !fusedH) ? p->d_lamHx : nullptr, st))) return rc;
    if (any_det && fusedH) {  // lambda_H_in += the detectors' H_prev cotangent (same addition, after the kernel)
      for (const auto& bx : boxes) {  // overlapping boxes: the first launch adds and clears the overlap, the next one sees zeros
        const int ex = bx[1] - bx[0], ey = bx[3] - bx[2], ez = bx[5] - bx[4];
        if (ex <= 0 || ey <= 0 || ez <= 0) continue;
        const long long nb = 3LL * ex * ey * ez;
        adj_box_add_clear_kernel<<<(unsigned)std::min<long long>((nb + 255) / 256, 148 * 16), 256, 0, st>>>(lamH, p->d_lamHx, p->nx, p->ny, p->nz, bx[0], bx[2], bx[4], ex,
                                                                                                          ey, ez);
        p->launches++;
        CUDA_TRY(cudaGetLastError());
      }
    }
    if (interleave_t) {
      if ((rc = fdtdx_b200_run_reverse_phase(p, t, 3, 0, 0, stream))) return rc;
      if ((rc = make_params(p, S, 1))) return rc;
      E = S.E;
      H = S.H;
    }
    // (5) E half-step transpose: lambda_E' -> lambda_E_in, accumulates into lambda_H_in
    if ((rc = adjoint_half(p, S, true, E, H, lamE, lamH, nullptr, st))) return rc;
  }
  return adj_cotpsi_home(p, st);
}

extern "C" int fdtdx_b200_run_adjoint_exact(FdtdxPlan* p, int t, void* stream) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  p->adjoint_exact = true;
  const int rc = fdtdx_b200_run_adjoint(p, t + 1, 1, stream);
  p->adjoint_exact = false;
  return rc;
}

extern "C" int fdtdx_b200_run_forward_host(FdtdxPlan* p, const float* h_E, const float* h_H, const float* h_inv_eps,
                                           float* h_E_out, float* h_H_out, int t0, int n, int record_detectors,
                                           void* stream, size_t* h2d_bytes, size_t* d2h_bytes) {
  if (!p) return fail(FDTDX_EINVAL, "null plan");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)p->nx * p->ny * p->nz;
  float* dE = (float*)p->slots[FDTDX_SLOT_E][0];
  float* dH = (float*)p->slots[FDTDX_SLOT_H][0];
  float* dEps = (float*)p->slots[FDTDX_SLOT_INV_EPS][0];
  if (!dE || !dH || !dEps) return fail(FDTDX_EUNBOUND, "E, H and INV_EPS must be bound");
  size_t up = 0, down = 0;
  if (h_E) { CUDA_TRY(cudaMemcpyAsync(dE, h_E, 3 * N * 4, cudaMemcpyHostToDevice, st)); up += 3 * N * 4; }
  if (h_H) { CUDA_TRY(cudaMemcpyAsync(dH, h_H, 3 * N * 4, cudaMemcpyHostToDevice, st)); up += 3 * N * 4; }
  if (h_inv_eps) {
    const size_t nb = (size_t)p->eps_tier * N * 4;
    CUDA_TRY(cudaMemcpyAsync(dEps, h_inv_eps, nb, cudaMemcpyHostToDevice, st));
    up += nb;
  }
  int rc = fdtdx_b200_run_forward(p, t0, n, record_detectors, 0, 1, stream);
  if (rc) return rc;
  if (h_E_out) { CUDA_TRY(cudaMemcpyAsync(h_E_out, dE, 3 * N * 4, cudaMemcpyDeviceToHost, st)); down += 3 * N * 4; }
  if (h_H_out) { CUDA_TRY(cudaMemcpyAsync(h_H_out, dH, 3 * N * 4, cudaMemcpyDeviceToHost, st)); down += 3 * N * 4; }
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h2d_bytes) *h2d_bytes = up;
  if (d2h_bytes) *d2h_bytes = down;
  return FDTDX_OK;
}
