// forces_nn.cu -- translation unit of the multiphase non-Newtonian force kernels (forces_nn.cuh).
#include "forces_api.h"
#include "forces_nn.cuh"
