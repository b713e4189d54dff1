// aes_common.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/aesim.h"

#define AES_EXPORT extern "C" __attribute__((visibility("default")))

void aes_set_error(const char *fmt, ...);
void aes_count_launch();

#define AES_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            aes_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),        \
                          __FILE__, __LINE__);                                           \
            return AES_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define AES_REQUIRE(cond, ...)                                                           \
    do {                                                                                 \
        if (!(cond)) { aes_set_error(__VA_ARGS__); return AES_ERR_INVALID; }             \
    } while (0)
