// Library identification for the C ABI (include/densefusion_b200.h).
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

#ifndef DF_HAVE_TC
#define DF_HAVE_TC 0
#endif

extern "C" int df_abi_version(void) { return 1; }
extern "C" int df_features(void) { return DF_HAVE_TC ? 1 : 0; }
