// C-ABI entry points of liblivecc_sm100a.so (declared in include/livecc_b200.h).
// Plain pointers and sizes only; never throws; every call is asynchronous on the given stream.
#include "../../include/livecc_b200.h"
#include "cabi_common.h"
#include "gemm.h"

extern "C" {

int lcc_abi_version(void) { return LCC_ABI_VERSION; }

lcc_ctx* lcc_create(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
    if (prop.major != 10) return nullptr;  // sm_100a only: there is no fallback path
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    lcc_ctx* c = new lcc_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->err[0] = 0;
    return c;
}

void lcc_destroy(lcc_ctx* ctx) { delete ctx; }

const char* lcc_last_error(lcc_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

int lcc_num_sms(lcc_ctx* ctx) { return ctx ? ctx->num_sms : -1; }

int lcc_gemm_bf16(lcc_ctx* ctx, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                  int M, int N, int K, const void* bias, const void* residual, int ldr, int epilogue,
                  int block_n, lcc_stream_t stream) {
    if (!ctx) return -1;
    lcc::GemmArgs a;
    a.A = A; a.B = B; a.C = C;
    a.M = M; a.N = N; a.K = K;
    a.lda = lda; a.ldb = ldb; a.ldc = ldc;
    a.bias = bias; a.residual = residual; a.ldr = ldr;
    a.epi = epilogue; a.block_n = block_n;
    int r = lcc::gemm_bf16_tn(a, ctx->num_sms, (cudaStream_t)stream);
    if (r) LCC_FAIL(ctx, r, "lcc_gemm_bf16 failed (code %d; M=%d N=%d K=%d epi=%d)", r, M, N, K, epilogue);
    return 0;
}

}  // extern "C"
