/* tests/emu/stub/cuda_runtime.h -- stands in for the CUDA runtime header when a kernel header is
 * compiled for the HOST by tests/emu (see cuda_emu.h).  Test infrastructure only. */
#pragma once
#include "../cuda_emu.h"
