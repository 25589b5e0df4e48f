// gdm/system.h -- same header name as the reference's include/gdm/system.h; the B200-native
// implementation lives in gdm/gdm.h (C++ front end over the C ABI gdm/cuda/gdm_c_api.h).
#pragma once
#include "gdm.h"
