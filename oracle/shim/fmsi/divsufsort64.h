/* Build shim, see divsufsort.h in this directory. Declaration only; never called by FMSI. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int64_t saidx64_t;
int32_t divsufsort64(const uint8_t *T, saidx64_t *SA, saidx64_t n);
#ifdef __cplusplus
}
#endif
