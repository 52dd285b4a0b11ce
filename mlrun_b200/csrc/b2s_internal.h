// b2s_internal.h -- what the translation units of libb200serve.so share (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>

#define B2S_HIDDEN __attribute__((visibility("hidden")))

B2S_HIDDEN int b2s_int_fail(int code, const char* fmt, ...);  // sets b2s_last_error(), returns code
B2S_HIDDEN bool b2s_int_inited();
B2S_HIDDEN int b2s_int_device();
B2S_HIDDEN int b2s_int_sm_count();
B2S_HIDDEN cudaStream_t b2s_int_stream();                      // the library stream
B2S_HIDDEN void b2s_int_count_launches(int n);
struct b2s_plan_s;
B2S_HIDDEN int b2s_int_plan_shape(b2s_plan_s* plan, int* n_in, int* out_cols);  // B2S_ERR_STATE unless finalized
