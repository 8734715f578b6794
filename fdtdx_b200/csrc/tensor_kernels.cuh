// Full-tensor (9-component) material path - kernels. (filled in below)
#pragma once
#include "common.cuh"
