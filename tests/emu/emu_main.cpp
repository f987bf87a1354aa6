// TEST INFRASTRUCTURE — instantiates the emulator globals for libmvmc_emu.so (see cuda_emu.h).
#define MVMC_EMU_IMPL
#include "cuda_emu.h"
