// TMA-staged E half-step with flat tiles (grids whose rows are neither 64 nor 128 cells; see TmaRt in yee_tma.cuh).
#define FDTDX_TZ_SEL 0
#define fdtdx_dispatch_E4_tma fdtdx_dispatch_E4_tmaf
#include "yee_E4t.cu"
