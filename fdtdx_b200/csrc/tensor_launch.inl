// host-side launcher of the full-tensor tier (included by abi.cu after make_params is defined)
static int tensor_step(FdtdxPlan* p, int t, int simulate, bool rev, bool is_E, cudaStream_t st) {
  if (p->nx != p->nxg) return fail(FDTDX_EUNSUPPORTED, "full-tensor media are not supported on x-sharded plans");
  StepParams S;
  int rc = make_params(p, S, simulate);
  if (rc) return rc;
  const long long N = (long long)p->nx * p->ny * p->nz;
  if (!p->d_K) {
    rc = to_device<float>(p, nullptr, (size_t)3 * N, &p->d_K);
    if (rc) return rc;
  }
  TensorParams T;
  memset(&T, 0, sizeof(T));
  T.nx = p->nx; T.ny = p->ny; T.nz = p->nz;
  for (int a = 0; a < 3; ++a) { T.wrap[a] = p->wrap[a]; T.pml[a] = S.pml[a]; T.w[a] = p->d_w[a]; }
  T.cour = S.cour; T.dt = S.dt;
  float* E0 = (float*)p->slots[FDTDX_SLOT_E][0];
  float* E1 = (float*)p->slots[FDTDX_SLOT_E_ALT][0];
  float* H0 = (float*)p->slots[FDTDX_SLOT_H][0];
  float* H1 = (float*)p->slots[FDTDX_SLOT_H_ALT][0];
  float* Ecur = p->e_parity ? E1 : E0;
  float* Hcur = p->h_parity ? H1 : H0;
  if (is_E) {
    if (!E1) return fail(FDTDX_EUNBOUND, "E_ALT must be bound for the full-tensor E update");
    T.F_in = Ecur; T.F_out = p->e_parity ? E0 : E1; T.F_other = Hcur;
    T.A = (const float*)p->slots[FDTDX_SLOT_TENSOR_A_E][0];
    T.B = (const float*)p->slots[FDTDX_SLOT_TENSOR_B_E][0];
    T.mat = S.eps; T.mat_tier = p->eps_tier; T.mat_cs = (p->eps_tier == 1) ? 0 : N; T.mat_scalar = 1.0f;
    for (int a = 0; a < 3; ++a) T.sc[a] = p->d_sB[a];
  } else {
    if (!H1) return fail(FDTDX_EUNBOUND, "H_ALT must be bound for the full-tensor H update");
    T.F_in = Hcur; T.F_out = p->h_parity ? H0 : H1; T.F_other = Ecur;
    T.A = (const float*)p->slots[FDTDX_SLOT_TENSOR_A_H][0];
    T.B = (const float*)p->slots[FDTDX_SLOT_TENSOR_B_H][0];
    T.mat = S.mu; T.mat_tier = p->mu_tier; T.mat_cs = (p->mu_tier <= 1) ? 0 : N; T.mat_scalar = S.inv_mu_scalar;
    for (int a = 0; a < 3; ++a) T.sc[a] = p->d_sF[a];
  }
  if (!T.B) return fail(FDTDX_EUNBOUND, "TENSOR_B must be bound for the full-tensor update");
  T.K = p->d_K;
  T.simulate = simulate; T.is_E = is_E ? 1 : 0; T.reverse = rev ? 1 : 0;
  T.n_walls = S.n_walls; T.walls = S.walls; T.n_src = S.n_src; T.src = S.src;
  if (is_E && p->n_poles > 0) {
    if (rev) return fail(FDTDX_EUNSUPPORTED, "Dispersive time-reversible gradient computation under active development. Use GradientConfig(method='checkpointed') instead.");
    if (p->has_c4) return fail(FDTDX_EUNSUPPORTED, "CCPR (c4) poles are rejected for the full-tensor branch (update.py:406)");
    T.n_poles = p->n_poles; T.P_cur = S.P_cur; T.P_new = S.P_new; T.c1 = S.c1; T.c2 = S.c2; T.c3 = S.c3; T.c_cs = S.c_cs;
  }
  const unsigned blocks = (unsigned)((N + 255) / 256);
  if (rev) {
    // sources first, in place on the current buffer
    for (size_t si = 0; si < p->srcs.size(); ++si) {
      const SrcDev& d = p->srcs[si].d;
      const long long n = (long long)(d.hi[0] - d.lo[0]) * (d.hi[1] - d.lo[1]) * (d.hi[2] - d.lo[2]);
      if (n <= 0) continue;
      TensorParams R = T;
      R.F_out = const_cast<float*>(T.F_in);
      tensor_inject_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(R, t, (int)si);
      p->launches++;
    }
  }
  // 128-bit fast path: phase 1 = the marching half-step kernel in curl-only mode (register queue,
  // shuffles, L2 prefetch), phase 2 = tensor_apply4_kernel, sources by O(surface) launches afterwards
  bool fast = (p->nz % 4 == 0) && can_vec4(p, S) && aligned16(T.K) && aligned16(T.A) && aligned16(T.B) && aligned16(T.F_in) &&
              aligned16(T.F_out) && aligned16(T.F_other);
  {
    const char* e = getenv("FDTDX_B200_TENSOR_FAST");
    if (e && e[0] == '0') fast = false;
  }
  if (fast && T.n_poles == 0) {
    StepParams Q = S;
    if (is_E) { Q.H = const_cast<float*>(T.F_other); Q.E = T.K; }
    else { Q.E = const_cast<float*>(T.F_other); Q.H = T.K; }
    Q.n_src = 0;
    if (can_tma(p, Q, true) && tma_tz(p) == 128) {
      // TMA-staged curl-only kernel: three halo tiles of the other field per plane, K stored with 128-bit writes
      TmaSet M;
      memset(&M, 0, sizeof(M));
      if ((rc = get_tmap(p, T.F_other, 3, p->nx, 0, &M.fld_halo))) return rc;
      M.fld_plain = M.mat_plain = M.xhalo = M.fld_halo;  // not read in curl-only mode
      Q.xchunk = tma_chunk(p, Q);
      dim3 g((p->nz + 127) / 128, (p->ny + FDTDX_TMA_R - 1) / FDTDX_TMA_R, (Q.x_end - Q.x_begin + Q.xchunk - 1) / Q.xchunk);
      if (is_E) CUDA_TRY(fdtdx_dispatch_E4_tma_konly(Q, M, t, pml_mode(p, Q), rev, p->metric, g, st));
      else CUDA_TRY(fdtdx_dispatch_H4_tma_konly(Q, M, t, pml_mode(p, Q), rev, p->metric, g, st));
    } else {
      dim3 b(32, p->rows);
      dim3 g((p->nz + 127) / 128, (p->ny + p->rows - 1) / p->rows, (Q.x_end - Q.x_begin + Q.xchunk - 1) / Q.xchunk);
      if (is_E) fdtdx_dispatch_E4_konly(Q, t, pml_mode(p, Q), rev, p->metric, g, b, st);
      else fdtdx_dispatch_H4_konly(Q, t, pml_mode(p, Q), rev, p->metric, g, b, st);
    }
  } else {
    tensor_curl_kernel<<<blocks, 256, 0, st>>>(T);
  }
  if (fast && T.w[0] == nullptr) {
    TensorParams T2 = T;
    T2.n_src = 0;
    dim3 b(32, 8), g((p->nz + 127) / 128, (p->ny + 7) / 8, p->nx);
    if (is_E) { if (T2.A) tensor_apply4_kernel<true, true><<<g, b, 0, st>>>(T2); else tensor_apply4_kernel<true, false><<<g, b, 0, st>>>(T2); }
    else { if (T2.A) tensor_apply4_kernel<false, true><<<g, b, 0, st>>>(T2); else tensor_apply4_kernel<false, false><<<g, b, 0, st>>>(T2); }
    if (!rev) {
      for (size_t si = 0; si < p->srcs.size(); ++si) {
        const SrcDev& d = p->srcs[si].d;
        const long long n = (long long)(d.hi[0] - d.lo[0]) * (d.hi[1] - d.lo[1]) * (d.hi[2] - d.lo[2]);
        if (n <= 0) continue;
        tensor_inject_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(T, t, (int)si, 0);
        p->launches++;
      }
    }
  } else {
    tensor_apply_kernel<<<blocks, 256, 0, st>>>(T, t);
  }
  p->launches += 2;
  CUDA_TRY(cudaGetLastError());
  if (is_E) {
    p->e_parity ^= 1;
    if (p->n_poles > 0) p->p_parity ^= 1;
  } else {
    p->h_parity ^= 1;
  }
  return FDTDX_OK;
}
