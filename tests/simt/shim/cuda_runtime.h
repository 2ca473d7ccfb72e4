// Stand-in for <cuda_runtime.h> when the kernel sources (and, for the whole-ABI build, diskann_b200.cu) are compiled
// by g++ for the SIMT emulator (tests/simt/simt_emu.h).  Test infrastructure only.
#pragma once
#include "../simt_emu.h"
#include "fake_cuda.h"
