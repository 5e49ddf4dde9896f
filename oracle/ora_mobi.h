/* placeholder, replaced below */
#ifndef UVIC_ORA_MOBI_H
#define UVIC_ORA_MOBI_H
#define ORA_MOBI_NIDX 128
struct ora_mobi_par { double p[256]; };
#endif
