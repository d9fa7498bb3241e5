/*
 * oracle/th_shim/TH/TH.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for the long-gone Torch7 `TH/TH.h`, just enough for the reference's CPU file
 * old_GEOMetrics/chamfer_distance/src/my_lib.c (which uses ->size[k], THFloatTensor_data and
 * THIntTensor_data and nothing else) to compile UNMODIFIED from /root/reference into
 * oracle/_ref/libref_cpu.so.  ctypes mirrors of these structs live in oracle/refload.py.
 */
#ifndef GEOM_ORACLE_TH_SHIM_H
#define GEOM_ORACLE_TH_SHIM_H

typedef struct THFloatTensor { long size[4]; float *data; } THFloatTensor;
typedef struct THIntTensor { long size[4]; int *data; } THIntTensor;

static inline float *THFloatTensor_data(THFloatTensor *t) { return t->data; }
static inline int *THIntTensor_data(THIntTensor *t) { return t->data; }

#endif
