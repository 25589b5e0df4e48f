// gdm/fe.h -- same header name as the reference's include/gdm/fe.h; the B200-native
// implementation lives in gdm/gdm.h (C++ front end over the C ABI gdm/cuda/gdm_c_api.h).
#pragma once
#include "gdm.h"
