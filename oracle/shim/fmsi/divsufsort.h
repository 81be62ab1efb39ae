/* Build shim (test infrastructure, not product code).
 * The reference's main.cpp includes <sdsl/suffix_arrays.hpp>, which includes "divsufsort.h".
 * Upstream generates that header with cmake from divsufsort.h.cmake; we do not run the
 * reference's build system (oracle/Makefile compiles its sources directly), so this file
 * only DECLARES the one entry point sdsl's construct_sa.hpp names. FMSI never calls it
 * (suffix sorting goes through QSufSort.c), so no definition is linked. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef uint8_t sauchar_t;
typedef int32_t saint_t;
typedef int32_t saidx_t;
saint_t divsufsort(const sauchar_t *T, saidx_t *SA, saidx_t n);
#ifdef __cplusplus
}
#endif
