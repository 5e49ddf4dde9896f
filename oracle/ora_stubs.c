#include "oracle.h"
void ora_mobi_columns(ora_ctx *c) { (void)c; }
void ora_filt(ora_ctx *c) { (void)c; }
