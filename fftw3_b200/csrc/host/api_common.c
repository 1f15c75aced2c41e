/* api_common.c -- precision-independent B200 extensions declared at the end of
 * include/fftw3.h (stream selection, async mode, launch counter). */
#include "b2_internal.h"

void fftw_b200_set_stream(void *cuda_stream) { b2d_set_stream(cuda_stream); }
void fftw_b200_set_async(int enabled) { b2_async_mode = enabled ? 1 : 0; }
void fftw_b200_synchronize(void) { b2d_sync(); }
unsigned long long fftw_b200_launch_count(void) { return (unsigned long long)b2d_launch_count(); }
const char *fftw_b200_device_name(void) { return b2d_device_name(); }
