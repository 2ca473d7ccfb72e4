// Stand-in for CUDA's <math_constants.h> under the CPU SIMT emulator (nothing from it is used by the kernels).
#pragma once
