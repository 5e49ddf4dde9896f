/* placeholder until the Fourier filter restatement lands (source/common/filt.F) */
#include "oracle.h"
void ora_filt(ora_ctx *c) { (void)c; }
