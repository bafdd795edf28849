/*
 * linetr_b200 - debug / micro-benchmark hooks of the CUDA library.  NOT part of the drop-in
 * boundary (include/linetr_b200.h); used by tools/ and the engine unit tests only.
 */
#ifndef LINETR_B200_DEBUG_H_
#define LINETR_B200_DEBUG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Micro-benchmark of the image-operand GEMM engine (zero-filled operands, timing only):
 * average device milliseconds per launch.  out_mode: 0 fp32 rows, 1 image, 2 both. */
float ltr_gemm_bench(int32_t m, int32_t n, int32_t k, int32_t bn_hint, int32_t out_mode, int32_t iters,
                     int32_t device);

/* clock64 stamps of CTA 0 recorded by the last ltr_gemm_bench call (64 slots, 16 per tile:
 * 0 MMA acc_empty ok, 1 first operands landed, 2 MMAs issued, 3 epilogue acc_full ok, 4 first
 * TMEM chunk read, 5 epilogue done, 6 producer slot free). */
const unsigned long long* ltr_gemm_trace(void);

/* Kernels instrumented with LTR_DBG_STAMP store clock64 stamps of their first CTA into a
 * 128-slot device array while tracing is armed. */
void ltr_debug_trace_arm(int32_t on);
int ltr_debug_trace_read(unsigned long long* out128);

#ifdef __cplusplus
}
#endif
#endif /* LINETR_B200_DEBUG_H_ */
