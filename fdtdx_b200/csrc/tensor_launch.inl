// host-side launcher of the full-tensor path (included by abi.cu)
static int tensor_step(FdtdxPlan* p, int t, int simulate, bool rev, bool is_E, cudaStream_t st) {
  (void)p; (void)t; (void)simulate; (void)rev; (void)is_E; (void)st;
  return fail(FDTDX_EUNSUPPORTED, "full-tensor material path not built yet");
}
