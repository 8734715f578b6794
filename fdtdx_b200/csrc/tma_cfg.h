// Compile-time geometry of the TMA-staged kernels, shared by the kernels' dispatchers and the
// tensor-map builder in abi.cu (the tile box is baked into the tensor maps).
#pragma once
#ifndef FDTDX_TMA_R
#define FDTDX_TMA_R 8  // y rows (consumer warps) per CTA
#endif
#ifndef FDTDX_TMA_S
#define FDTDX_TMA_S 3  // ring depth (planes in flight per CTA)
#endif
