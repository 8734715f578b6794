// TMA-staged H half-step for thin grids (Nz <= 64): 64-cell tile rows, two rows per warp (see yee_tma.cuh).
#define FDTDX_TZ_SEL 64
#define fdtdx_dispatch_H4_tma fdtdx_dispatch_H4_tma64
#include "yee_H4t.cu"
