/*
 * liblshm_sm100 - self-test hooks.  NOT part of the product ABI (include/lshm.h): host-only helpers
 * that let the CPU test-suite exercise pieces of host-side index math the kernels rely on.
 */
#ifndef LSHM_SELFTEST_H_
#define LSHM_SELFTEST_H_

#include "lshm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Multiply-shift division used by the conv kernels for their index math (conv_geom.cuh FastDiv:
 * q = (mulhi(m, n) + n) >> l): *mismatches = how many of the `count` 31-bit values n[i] give a quotient
 * different from n[i] / d.  Host pointers, no GPU needed. */
LSHM_API int lshm_fastdiv_check(int64_t d, const int64_t* n, int count, int64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* LSHM_SELFTEST_H_ */
