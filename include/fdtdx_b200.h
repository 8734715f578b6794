/*
 * fdtdx_b200.h - C ABI of the B200-native (sm_100a) Yee time-stepping backend.
 *
 * This is the drop-in boundary for fdtdx's hot path.  The reference has no
 * FFI today: its seams are the Python call sites
 *     run_fdtd              src/fdtdx/fdtd/wrapper.py:14-63
 *     checkpointed_fdtd     src/fdtdx/fdtd/fdtd.py:421-496
 *     custom_fdtd_forward   src/fdtdx/fdtd/fdtd.py:499-584
 *     reversible_fdtd       src/fdtdx/fdtd/fdtd.py:39-418   (custom_vjp: fdtd_fwd/fdtd_bwd)
 *     forward / backward    src/fdtdx/fdtd/forward.py:83-156, backward.py:18-135
 * Each entry point below replaces the loop body those drivers run (the citation on each function
 * names the reference code it stands in for).  INTEGRATION.md shows the jax.ffi / ctypes stubs a
 * maintainer would add on the reference side.
 *
 * Conventions: extern "C"; plain pointers and sizes; no exceptions cross the boundary - every call
 * returns 0 on success or a negative FDTDX_E* code, with fdtdx_b200_last_error() giving the text.
 * The caller owns every field / material / state buffer (device memory, reference layouts:
 * (3,Nx,Ny,Nz) float32 C-order, z fastest).  A plan owns only small constant tables and scratch.
 * All work is enqueued on the caller's cudaStream_t (passed as void*).  A plan is bound to the
 * device that was current at fdtdx_b200_plan_create and is not re-entrant.
 */
#ifndef FDTDX_B200_H
#define FDTDX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct FdtdxPlan FdtdxPlan;

#define FDTDX_OK 0
#define FDTDX_EINVAL (-1)
#define FDTDX_ECUDA (-2)
#define FDTDX_EUNBOUND (-3)
#define FDTDX_EUNSUPPORTED (-4)

/* Buffer slots for fdtdx_b200_bind (device pointers, caller-owned). `index` selects the
 * sub-buffer where a slot holds several (documented per slot). */
enum FdtdxSlot {
  FDTDX_SLOT_E = 0,        /* (3,nx,ny,nz) f32  - FieldState.E           container.py:372 */
  FDTDX_SLOT_H = 1,        /* (3,nx,ny,nz) f32  - FieldState.H           container.py:375 */
  FDTDX_SLOT_INV_EPS = 2,  /* (tier,nx,ny,nz)   - inv_permittivities     container.py:405 */
  FDTDX_SLOT_INV_MU = 3,   /* (tier,nx,ny,nz) or unbound when scalar     container.py:408 */
  FDTDX_SLOT_SIGMA_E = 4,  /* (tier,...) pre-scaled conductivity         container.py:417 */
  FDTDX_SLOT_SIGMA_H = 5,
  FDTDX_SLOT_PSI_E = 6,    /* index = 2*pml + {0,1}; slab-shaped f32     container.py:378-381 */
  FDTDX_SLOT_PSI_H = 7,
  FDTDX_SLOT_P_A = 8,      /* (np,3,nx,ny,nz) dispersive_P_curr at even parity */
  FDTDX_SLOT_P_B = 9,      /* (np,3,nx,ny,nz) dispersive_P_prev at even parity */
  FDTDX_SLOT_C1 = 10,
  FDTDX_SLOT_C2 = 11,
  FDTDX_SLOT_C3 = 12,
  FDTDX_SLOT_C4 = 13,
  FDTDX_SLOT_DET_STATE = 14, /* index = 4*detector + key  (key order: see fdtdx_b200_plan_add_detector) */
  FDTDX_SLOT_REC_DATA = 15,  /* index = 2*pml + {0:E,1:H}; (slots,3,*face) of the recorder dtype */
  FDTDX_SLOT_E_ALT = 16,     /* full-tensor path: second E buffer (ping-pong) */
  FDTDX_SLOT_H_ALT = 17,
  FDTDX_SLOT_TENSOR_A_E = 18, /* (9,nx,ny,nz) precomputed A (NULL => identity) fdtd/misc.py:69-98 */
  FDTDX_SLOT_TENSOR_B_E = 19, /* (9,nx,ny,nz) precomputed B */
  FDTDX_SLOT_TENSOR_A_H = 20,
  FDTDX_SLOT_TENSOR_B_H = 21,
  FDTDX_SLOT_HALO_H_LO = 22,  /* (2,ny,nz): Hy,Hz of plane x0-1 (x-slab neighbour), section 8e */
  FDTDX_SLOT_HALO_E_HI = 23,  /* (2,ny,nz): Ey,Ez of plane x1 */
  FDTDX_SLOT_GRAD_INV_EPS = 24, /* adjoint: gradient accumulator, same shape as inv_eps */
  FDTDX_SLOT_GRAD_INV_MU = 25,
  FDTDX_SLOT_COT_E = 26,      /* adjoint carry lambda_E (3,nx,ny,nz) */
  FDTDX_SLOT_COT_H = 27,
  FDTDX_SLOT_COT_PSI_E = 28,  /* index as PSI_E */
  FDTDX_SLOT_COT_PSI_H = 29,
  FDTDX_SLOT_COT_DET = 30,    /* index as DET_STATE: cotangent of the detector state */
  /* ADE adjoint (run_adjoint_exact on dispersive plans): cotangents of dispersive_P_curr / _prev, carried
   * in place like COT_E / COT_H, and optional gradient accumulators shaped like c1..c4 */
  FDTDX_SLOT_COT_P = 31,
  FDTDX_SLOT_COT_P_PREV = 32,
  FDTDX_SLOT_GRAD_C1 = 33,
  FDTDX_SLOT_GRAD_C2 = 34,
  FDTDX_SLOT_GRAD_C3 = 35,
  FDTDX_SLOT_GRAD_C4 = 36,
  FDTDX_SLOT_BLOCH_E = 37,   /* complex (Bloch k != 0) runs: E of the partner system (Re <-> Im), bloch.py:61-96 */
  FDTDX_SLOT_BLOCH_H = 38,
  /* x-sharded plans whose detectors straddle the low slab edge: plane x0-1 of the lower neighbour,
   * (3,ny,nz) each - E and H after this step's updates, and H before its H update (the co-location
   * stencil of curl.py:86-224 reads x-1) */
  FDTDX_SLOT_DET_XLO_E = 39,
  FDTDX_SLOT_DET_XLO_H = 40,
  FDTDX_SLOT_DET_XLO_HPREV = 41,
  FDTDX_SLOT_COUNT = 42
};

enum FdtdxBoundaryKind { FDTDX_WALL_PEC = 0, FDTDX_WALL_PMC = 1 };
enum FdtdxProfileKind { FDTDX_PROFILE_CW = 0, FDTDX_PROFILE_PULSE = 1, FDTDX_PROFILE_TABLE = 2 };
enum FdtdxDetectorKind {
  FDTDX_DET_FIELD = 0,   /* objects/detectors/field.py:31-63         key0 "fields" */
  FDTDX_DET_ENERGY = 1,  /* objects/detectors/energy.py:90-143       key0 "energy" | keys 0..2 XY,XZ,YZ planes */
  FDTDX_DET_POYNTING = 2,/* objects/detectors/poynting_flux.py:171-195  key0 "poynting_flux" */
  FDTDX_DET_PHASOR = 3   /* objects/detectors/phasor.py:190-235      key0 "phasor" (complex64) */
};
enum FdtdxRecDtype { FDTDX_REC_F32 = 0, FDTDX_REC_BF16 = 1, FDTDX_REC_F16 = 2, FDTDX_REC_F8E4M3FNUZ = 3,
                     FDTDX_REC_F8E4M3FN = 4, FDTDX_REC_F8E5M2 = 5 };

const char* fdtdx_b200_last_error(void);
int fdtdx_b200_version(void);

/* ---- plan construction (host tables are copied; pointers need not outlive the call) ----------
 * Replaces the per-run constants the reference closes over when it traces `forward`
 * (config.courant_number config.py:165-177, metric scales core/physics/curl.py:10-39).
 * nx is the LOCAL x extent of this rank's slab, [x_offset, x_offset+nx) of nx_global.
 * sB/sF: per-axis backward/forward metric scales of the LOCAL slab (length nx/ny/nz) or NULL on a
 * uniform grid.  widths: per-axis cell widths (global x) for the detector co-location weights
 * (curl.py:42-83) or NULL.  inv_mu_scalar is used when INV_MU is not bound (python float 1.0,
 * initialization.py:752-753).  eps_tier/mu_tier in {1,3,9}; sigma tiers in {0,1,3,9}; mu_tier 0 = scalar. */
int fdtdx_b200_plan_create(FdtdxPlan** out, int nx, int ny, int nz, int x_offset, int nx_global,
                           double courant_number, double dt, int total_time_steps,
                           int eps_tier, int mu_tier, int sigma_e_tier, int sigma_h_tier,
                           double inv_mu_scalar, const int wrap[3],
                           const float* const sB[3], const float* const sF[3],
                           const float* const widths[3]);
int fdtdx_b200_plan_destroy(FdtdxPlan* plan);

/* CPML slab (perfectly_matched_layer.py:97-190): axis, direction (0 '-', 1 '+'), global slab
 * [lo,hi) along axis (full cross-section), six coefficient tables of length hi-lo:
 * a/b/(1/kappa) for the E-side (used by curl_H) and the H-side (used by curl_E). Returns index. */
int fdtdx_b200_plan_add_pml(FdtdxPlan* plan, int axis, int direction, int lo, int hi,
                            const float* a_E, const float* b_E, const float* inv_kappa_E,
                            const float* a_H, const float* b_H, const float* inv_kappa_H,
                            int kappa_is_one);
/* A slab may be registered as a superset [lo,hi) of the reference's slab whose extra cells carry zero
 * coefficients (a = b = 0, 1/kappa = 1: no correction) - the z-padded layout of ragged grids aligns slab starts
 * that way.  This names the reference's own range [lo_true,hi_true) of slab `index`, which still positions the
 * recorder's interface plane (boundary.py:117-144) and the field reset of the reversed pass (backward.py:117-122). */
int fdtdx_b200_plan_pml_set_true_range(FdtdxPlan* plan, int index, int lo_true, int hi_true);
/* PEC / PMC wall (pec.py:70-77, pmc.py:63-76): zero the two tangential comps on box [lo,hi). */
int fdtdx_b200_plan_add_wall(FdtdxPlan* plan, int kind, int axis, const int lo[3], const int hi[3]);

/* TFSF plane source (tfsf.py:193-409, 739-806).  Arrays (3,*face) float32 host.  sign = +1/-1
 * (direction).  cE/cH = courant * metric scale at the plane (backward/forward).  profile params:
 *   CW:    p0 = period, p1 = phase (wave_character.phase_shift), p2 = profile phase_shift, p3 = startup_time
 *   PULSE: p0 = 2*pi*f_c, p1 = phase, p2 = center phase_shift, p3 = t0 (6 sigma_t), p4 = 2*sigma_t^2
 *   TABLE: p0 = start_time, p1 = table dt, p2 = outside_value, p3 = nearest(1)/linear(0); table in `signal`
 * on/t_adj: per-time-step gate and remapped step (source.py:41-49) or NULL for the default switch.
 * h_filter: optional filtered-H table (tfsf.py:259-264). */
int fdtdx_b200_plan_add_plane_source(FdtdxPlan* plan, const int lo[3], const int hi[3], int normal_axis,
                                     int sign, const float* E_inc, const float* H_inc,
                                     const float* toff_E, const float* toff_H, int profile_kind,
                                     const double params[8], const float* signal, int signal_len,
                                     double static_amplitude, double cE, double cH,
                                     const uint8_t* on, const float* t_adj,
                                     const float* h_filter, int h_filter_len);
/* Complex (lossy-mode) plane source (objects/sources/mode.py:212-222, tfsf.py:266-283, 366-383): the
 * imaginary parts of the incident profile, injected in quadrature - incident = Re * amp(phase) +
 * Im * amp(phase - pi/2).  quadrature_phase = wave_character.phase_shift - pi/2 (rounded to float32
 * like the reference's weak scalar).  Arrays are (3, *face) like E_inc / H_inc. */
int fdtdx_b200_plan_source_set_quadrature(FdtdxPlan* plan, int source_index, const float* E_inc_imag,
                                          const float* H_inc_imag, double quadrature_phase);
/* Point dipole (dipole.py:195-277): scale = courant*amplitude*static. electric!=0 -> E update. */
int fdtdx_b200_plan_add_dipole(FdtdxPlan* plan, const int cell[3], int polarization, int electric,
                               double scale, int profile_kind, const double params[8],
                               const float* signal, int signal_len, const uint8_t* on, const float* t_adj);

/* Detector (detector.py:195-244 + per-type update).  on[t], arr_idx[t] tables of length T.
 * flags: bit0 exact_interpolation, bit1 inverse, bit2 reduce_volume, bit3 as_slices,
 *        bit4 slices-use-mean, bit5 keep_all_components, bit6 negative direction,
 *        bit7 large region (hint): accumulate with the row-marching 128-bit kernels; same results
 *        (slice means: same sums in a different, fixed order).  Ignored when rows are not 16-byte aligned.
 * comp_mask: bit c set = component c of (Ex,Ey,Ez,Hx,Hy,Hz) recorded (field / phasor).
 * weights: cell-volume (field/energy/phasor reduce) or face-area (Poynting reduce) weights of the
 * region, or NULL.  phasor_table: (T, nf) complex64 = exp(i*omega_f*t*dt) (phasor.py:221-224),
 * window: (T) float32, scale: static scale.  slice_idx: energy slice positions. aux = Poynting axis.
 * State keys (index for FDTDX_SLOT_DET_STATE): field {0}, energy {0} or {0:XY,1:XZ,2:YZ},
 * poynting {0}, phasor {0}. */
int fdtdx_b200_plan_add_detector(FdtdxPlan* plan, int kind, const int lo[3], const int hi[3], int flags,
                                 int comp_mask, int aux, const uint8_t* on, const int32_t* arr_idx,
                                 const float* weights, int n_freq, const float* phasor_table,
                                 const float* window, double scale, const int slice_idx[3]);

/* Recorder (interfaces/recorder.py:70-199, time_filter.py:139-257, modules.py:98-161):
 * slot_of_time[t] (-1 = not stored), replay_a/b/w[t] for reconstruction; dtype enum FdtdxRecDtype. */
/* x-sharded plans: the detector's region is this rank's part only, but a weighted mean (reduce_volume of
 * field / phasor detectors, detector.py:108) divides by the weight sum of the WHOLE region. */
int fdtdx_b200_plan_detector_set_wsum(FdtdxPlan* plan, int detector_index, double weight_sum);
int fdtdx_b200_plan_set_recorder(FdtdxPlan* plan, int dtype, int n_slots, const int32_t* slot_of_time,
                                 const int32_t* replay_a, const int32_t* replay_b, const float* replay_w);
/* ADE dispersion (update.py:316-350): n_poles, coefficient component tier (1|3), c4 present. */
int fdtdx_b200_plan_set_dispersion(FdtdxPlan* plan, int n_poles, int coeff_tier, int has_c4);
/* x-slab neighbours: 0 = domain edge (zero / local wrap), 1 = halo buffer bound. */
int fdtdx_b200_halo_bind(FdtdxPlan* plan, int has_lo_neighbour, int has_hi_neighbour);
/* The plan's grid carries pad_cells extra z cells beyond the caller's Nz (kept at zero by walls; host side:
 * plan.py::_z_padding): interface-recorder planes (interfaces/recorder.py) keep the caller's Nz. */
int fdtdx_b200_set_z_padding(FdtdxPlan* plan, int pad_cells);
/* config.symmetry (fdtd/update.py:92-198): on a symmetric axis the min-side halo is never wrapped (:121-125);
 * where an electric symmetry wall sits on the min face, the detector co-location stencil reads the
 * parity-weighted mirror partner instead of the zero halo (pad_fields_with_symmetry_mirror, :139-198). */
int fdtdx_b200_set_symmetry(FdtdxPlan* plan, const int symmetric_axes[3], const int electric_wall_axes[3]);
/* BlochBoundary.apply_pad_correction (objects/boundaries/bloch.py:61-96) for a non-zero Bloch vector.  The
 * complex fields run as two real systems (Re, Im), one plan each; every wrapped ghost value mixes the
 * system's own field with its partner's (FDTDX_SLOT_BLOCH_E / _H): low side F[N-1] * conj(phase) =
 * self * cos[a] + partner * sin[a], high side F[0] * phase = self * cos[a] - partner * sin[a].  Pass
 * (cos, sin) of k_a * L_a for the Re system and (cos, -sin) for the Im system; enable = 0 turns it off. */
int fdtdx_b200_set_bloch(FdtdxPlan* plan, int enable, const double cos_kL[3], const double sin_kL[3]);

int fdtdx_b200_bind(FdtdxPlan* plan, int slot, int index, void* device_ptr);

/* ---- execution ------------------------------------------------------------------------------ */
/* n forward steps from time step t0: body of `forward` (fdtd/forward.py:83-156):
 * update_E (update.py:256), update_H (:689), [collect_interfaces :1140], [update_detector_states :1040]. */
int fdtdx_b200_run_forward(FdtdxPlan* plan, int t0, int n, int record_detectors, int record_boundaries,
                           int simulate_boundaries, void* stream);
/* Halves of one step, for callers that interleave the x-slab halo exchange (section 8e):
 * phase 0 = update_E, phase 1 = update_H, phase 2 = record + detectors. */
int fdtdx_b200_run_forward_phase(FdtdxPlan* plan, int t, int phase, int record_detectors,
                                 int record_boundaries, int simulate_boundaries, void* stream);
/* One half-step restricted to the x planes [x_begin, x_end) of this rank's slab (which: 0 = E,
 * 1 = H).  Lets an x-slab caller run the interior while the halo plane is in flight and the edge
 * chunk afterwards.  Phase 3 of run_forward_phase = detector H_prev gather only. */
int fdtdx_b200_run_half_range(FdtdxPlan* plan, int t, int which, int x_begin, int x_end,
                              int simulate_boundaries, void* stream);
int fdtdx_b200_get_xchunk(FdtdxPlan* plan);
/* n reverse steps entered from state t_from (first step reconstructs t_from-1): body of `backward`
 * (fdtd/backward.py:62-135): add_interfaces (update.py:1181), update_H_reverse (:856),
 * update_E_reverse (:526), [apply_field_reset], [inverse detectors]. */
int fdtdx_b200_run_reverse(FdtdxPlan* plan, int t_from, int n, int record_detectors, int reset_fields,
                           void* stream);
/* n iterations of the reversible_fdtd backward loop (fdtd/fdtd.py:215-251 body_fn): one reverse
 * step, then the VJP of one forward step at the reconstructed state, accumulating
 * GRAD_INV_EPS / GRAD_INV_MU and carrying COT_E/COT_H/COT_PSI_*. */
/* One phase of the reversed step t (backward.py:62-135), for callers that interleave two plans (the Re / Im
 * systems of a complex run): 0 replay the recorded interfaces, 1 copy H for the inverse detectors, 2
 * update_H_reverse, 3 update_E_reverse, 4 field reset in the PML (if reset_fields), 5 inverse detectors. */
int fdtdx_b200_run_reverse_phase(FdtdxPlan* plan, int t, int phase, int record_detectors, int reset_fields, void* stream);
int fdtdx_b200_run_adjoint(FdtdxPlan* plan, int t_from, int n, void* stream);
/* VJP of ONE forward step t at the state that is currently bound (E_t, H_t, psi_t), without the
 * time-reversed reconstruction: the building block of the checkpointed gradient
 * (fdtd/fdtd.py:482-493, kind="checkpointed": stored / recomputed states instead of reversed ones).
 * Same cotangent and gradient slots as run_adjoint. */
int fdtdx_b200_run_adjoint_exact(FdtdxPlan* plan, int t, void* stream);

/* Current ping-pong parities after the last run (0: A is current). Callers that own P_A/P_B or
 * E/E_ALT need them to tell which buffer holds the current state. */
int fdtdx_b200_get_parity(FdtdxPlan* plan, int* p_parity, int* e_parity, int* h_parity);
int fdtdx_b200_set_parity(FdtdxPlan* plan, int p_parity, int e_parity, int h_parity);

/* Total electromagnetic energy of the bound fields, sum over all local cells of
 * compute_energy (core/physics/metrics.py:15-67, diagonal tiers), written as one float to d_out (device).
 * This is the per-step global reduction of EnergyThresholdCondition (fdtd/stop_conditions.py:81-147);
 * on x-sharded plans the caller adds the ranks' values. */
int fdtdx_b200_total_energy(FdtdxPlan* plan, float* d_out, void* stream);

/* Kernel launches issued by this plan since creation (bench.py's gpu_launches claim). */
long long fdtdx_b200_launch_count(FdtdxPlan* plan);
/* Tuning knob: x-chunk length of the marching kernels (0 = auto). */
int fdtdx_b200_set_tuning(FdtdxPlan* plan, int xchunk, int rows_per_block);
/* Kernel-path selection for the E/H half-steps.  enable: 1 = TMA-staged shared-memory pipeline (default
 * where applicable: Nz % 4 == 0, 16-byte aligned buffers, non-periodic y/z), 0 = register-marching kernels,
 * -1 = follow the FDTDX_B200_TMA environment variable.  xchunk_tma: x planes per CTA (0 = heuristic).
 * Pure tuning: results are bit-identical on both paths.  (No reference counterpart: XLA picks its own fusion.) */
int fdtdx_b200_set_tma(FdtdxPlan* plan, int enable, int xchunk_tma);

/* ---- peer-memory halo over NVLink (replaces the XLA collective-permute of the x-sharded arrays,
 * fdtd/initialization.py:598-611, core/jax/sharding.py:160-176) -----------------------------------------
 * Instead of exchanging packed planes, a rank's half-step kernels read the neighbour's boundary plane in
 * place: Hy,Hz of the low neighbour's last x plane (E half-step, curl.py:360-361) and Ey,Ez of the high
 * neighbour's first x plane (H half-step, curl.py:273-274), through CUDA-IPC mappings of the neighbour's
 * own E / H arrays.  Ordering is kept by two stream-ordered progress counters per rank.
 *   peer_export: what = 0 (bound E array), 1 (bound H array), 2 (this rank's progress flags; allocated on
 *                first use).  Writes the 64-byte cudaIpcMemHandle_t of the enclosing allocation and the byte
 *                offset of the buffer inside it; the caller ships both to the neighbour process.
 *   peer_attach: side = 0 (low-x neighbour: pass ITS H export + flags export), 1 (high-x neighbour: ITS E
 *                export + flags export); nx_peer = the neighbour's local Nx.  Requires halo_bind for that
 *                side.  Once every bound side is attached, run_forward / run_forward_phase use the peer path
 *                and need no HALO_* buffers.  All ranks must issue the same sequence of half-steps. */
int fdtdx_b200_peer_export(FdtdxPlan* plan, int what, unsigned char* handle64, long long* offset);
int fdtdx_b200_peer_attach(FdtdxPlan* plan, int side, const unsigned char* field_handle64, long long field_offset,
                           const unsigned char* flags_handle64, long long flags_offset, int nx_peer);
/* 1 if an in-kernel neighbour wait gave up (a rank stalled or issued fewer half-steps; the results of
 * this rank are then invalid), else 0.  Synchronises the device. */
int fdtdx_b200_peer_status(FdtdxPlan* plan);
/* Back to exchanged halo buffers (HALO_H_LO / HALO_E_HI): used when a neighbour could not map this rank. */
int fdtdx_b200_peer_detach(FdtdxPlan* plan);

/* Host-buffer convenience used for end-to-end timing: copies E,H,inv_eps from HOST memory into the
 * bound device buffers, runs n forward steps, copies E,H back.  Sizes in bytes are returned. */
int fdtdx_b200_run_forward_host(FdtdxPlan* plan, const float* h_E, const float* h_H, const float* h_inv_eps,
                                float* h_E_out, float* h_H_out, int t0, int n, int record_detectors,
                                void* stream, size_t* h2d_bytes, size_t* d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* FDTDX_B200_H */
