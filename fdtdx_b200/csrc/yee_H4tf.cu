// TMA-staged H half-step with flat tiles (grids whose rows are neither 64 nor 128 cells; see TmaRt in yee_tma.cuh).
#define FDTDX_TZ_SEL 0
#define fdtdx_dispatch_H4_tma fdtdx_dispatch_H4_tmaf
#include "yee_H4t.cu"
