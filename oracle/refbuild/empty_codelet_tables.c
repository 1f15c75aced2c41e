/* Empty codelet solver tables for the codelet-less build of the reference.
 * The reference's generated codelets (dft/scalar/codelets/*.c etc.) are not
 * in its git tree; these four tables are what its conf.c files expect
 * (dft/conf.c:43, rdft/conf.c:57-59).  Oracle-side only. */
#include "kernel/ifftw.h"
extern const solvtab X(solvtab_dft_standard);
extern const solvtab X(solvtab_rdft_r2cf);
extern const solvtab X(solvtab_rdft_r2cb);
extern const solvtab X(solvtab_rdft_r2r);
const solvtab X(solvtab_dft_standard) = { SOLVTAB_END };
const solvtab X(solvtab_rdft_r2cf) = { SOLVTAB_END };
const solvtab X(solvtab_rdft_r2cb) = { SOLVTAB_END };
const solvtab X(solvtab_rdft_r2r) = { SOLVTAB_END };
