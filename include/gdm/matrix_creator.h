// gdm/matrix_creator.h -- same header name as the reference's include/gdm/matrix_creator.h; the B200-native
// implementation lives in gdm/gdm.h (C++ front end over the C ABI gdm/cuda/gdm_c_api.h).
#pragma once
#include "gdm.h"
