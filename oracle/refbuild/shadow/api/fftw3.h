/* Shadow of the reference's api/fftw3.h for building the reference's own
 * tests/bench harness against the PRODUCT library: redirects to our header so
 * the harness proves source compatibility of include/fftw3.h. */
#include "../../../../include/fftw3.h"
