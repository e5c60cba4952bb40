// Internal: context object behind the opaque lcc_ctx of include/livecc_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

struct lcc_ctx {
    int device;
    int num_sms;
    char err[512];
};

#define LCC_FAIL(ctx, code, ...)                               \
    do {                                                       \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return (code);                                         \
    } while (0)

#define LCC_CHECK_LAUNCH(ctx, what)                                                        \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) LCC_FAIL(ctx, -100, "%s: %s", what, cudaGetErrorString(e__)); \
    } while (0)
