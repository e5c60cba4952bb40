/*
 * livecc_b200.h — C ABI of liblivecc_sm100a.so, the B200-native (sm_100a) kernels behind LiveCC's
 * per-chunk encode -> prefill -> decode hot path.
 *
 * The reference (showlab/livecc) has no native/FFI boundary of its own: the path sits behind
 * Python surfaces (REF/demo/infer.py:43-50,165-174 -> transformers Qwen2VLForConditionalGeneration).
 * This header is therefore the boundary *we* define one level below that surface;
 * each entry point cites the reference interface (file:line) whose arithmetic it replaces.
 * `mq2vl.py` = transformers/models/qwen2_vl/modeling_qwen2_vl.py (transformers 5.5.0).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller (PyTorch storage in our host code).
 *   - Every call is asynchronous on `stream`; nothing synchronises the device.
 *   - Return value: 0 on success, negative on error; lcc_last_error(ctx) returns a message.
 *   - A ctx is bound to one device and one host thread at a time; no hidden globals.
 *   - bf16 everywhere unless the name says otherwise; "bf16 rounding points" follow the reference's
 *     eager graph (one torch op = one rounding).
 */
#ifndef LIVECC_B200_H
#define LIVECC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCC_ABI_VERSION 1

typedef struct lcc_ctx lcc_ctx;
typedef void* lcc_stream_t; /* cudaStream_t */

/* GEMM epilogues (lcc_gemm_bf16) */
#define LCC_EPI_NONE 0            /* C = bf16(acc) */
#define LCC_EPI_BIAS 1            /* C = bf16(acc + bias) */
#define LCC_EPI_BIAS_QUICKGELU 2  /* ViT fc1: mq2vl.py:329-337 */
#define LCC_EPI_BIAS_GELU 3       /* merger mlp[0..1]: mq2vl.py:317-326 */
#define LCC_EPI_RESIDUAL 4        /* o_proj / down_proj + residual: mq2vl.py:645-660 */
#define LCC_EPI_BIAS_RESIDUAL 5   /* ViT proj / fc2 + residual: mq2vl.py:479-487 */
#define LCC_EPI_SWIGLU 6          /* gate|up (16-row interleaved) -> silu(gate)*up: mq2vl.py:502-504 */

int lcc_abi_version(void);

/* Context. lcc_create returns NULL unless `device` is an sm_100 GPU (there is no fallback path). */
lcc_ctx* lcc_create(int device);
void lcc_destroy(lcc_ctx* ctx);
const char* lcc_last_error(lcc_ctx* ctx);
int lcc_num_sms(lcc_ctx* ctx);

/*
 * C[M,N] = epilogue(A[M,K] x B[N,K]^T), bf16 in, fp32 accumulate (tcgen05/TMEM), bf16 out.
 * Replaces every nn.Linear / Conv3d-as-GEMM call with M >= 16 on the path:
 * mq2vl.py:304-310 (patch embed), :385,:401-403,:456 (ViT qkv/proj), :329-337 (ViT MLP),
 * :317-326 (merger), :539-541,:559-565,:593 (decoder q/k/v/o), :502-504 (MLP), :1437-1438 (lm_head).
 * K, N and all leading dimensions must be multiples of 8 elements.
 * block_n: 0 = heuristic, else 64/128/256.
 */
int lcc_gemm_bf16(lcc_ctx* ctx, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                  int M, int N, int K, const void* bias, const void* residual, int ldr, int epilogue,
                  int block_n, lcc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIVECC_B200_H */
