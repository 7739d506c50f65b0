// stand-in for <cuda_runtime.h> in the TOE_EMU test build (see ../cuda_emu.h)
#pragma once
#include "../cuda_emu.h"
