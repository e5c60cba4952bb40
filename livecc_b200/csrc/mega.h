// Internal interface of the persistent decode-step kernel (decode_mega.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcc {

constexpr int MG_MAXB = 8;         // streams per launch (the 8 columns of an m16n8k16 B operand)
constexpr int MG_THREADS = 352;    // 8 consumer warps + 2 TMA producer warps + 1 L2 prefetch warp (one lane each)
constexpr int MG_MAX_SLOTS = 16;   // ring groups of 16 KB
constexpr int MG_MAX_ITEMS = 64;   // split-KV items per (stream, kv head)

struct MegaLayer {
    const __nv_bfloat16 *ln1_w, *qkv_b, *ln2_w;
};

struct MegaStream {
    const int* page_table;
    int* scalars;
};

struct MegaParams {
    const CUtensorMap* wmaps;  // device array [4*L + 1]: per layer qkv, o, gate_up, down; then lm_head
    const MegaLayer* layers;   // device array [L]
    const __nv_bfloat16* final_norm_w;
    const float* inv_freq;
    int L, H, I, Hq, Hkv, V, qkv_dim;
    float eps;
    int B;
    MegaStream st[MG_MAXB];
    __nv_bfloat16 *k_pool, *v_pool;
    long long layer_stride;   // elements between layers of a pool
    int kv_rows_per_layer;    // = num_pages * Hkv * 64 (rows of the 2-D [rows,128] view per layer)
    __nv_bfloat16 *h, *qkv, *attn, *act;  // [8][H], [8][qkv_dim], [8][Hq*128], [8][I]
    float *logits_raw, *logits_proc;      // [8][V]
    float *part_o, *part_ml;              // [8][Hkv][MG_MAX_ITEMS][8][128] / [...][2]
    int* pair_cnt;                        // [8*Hkv], zero between launches
    unsigned* bar;                        // grid barrier counter, zero at launch
    int* err;                             // sticky error flag (bounded waits)
    unsigned long long* trace;            // optional [grid][64] globaltimer stamps (bring-up/profiling aid), or null
    int ngroup;      // ring depth in 16 KB groups
    int lookahead;   // groups (16 KB) per CTA the L2 prefetcher may run ahead of the ring (0 = no prefetcher)
    int layer_begin, layer_end, phase_mask, do_head;  // sub-range execution (tests); full step = 0, L, 31, 1
    float scale_log2;
};

int mega_smem_bytes(int H, int B, int* ngroup_out);
int mega_make_weight_tmap(CUtensorMap* tm, const void* w, int N, int K);
// k_pool/v_pool: whole pools viewed as [pool_rows, 128] bf16. Enqueues the kernel on `s` (p.bar must be zero).
int decode_mega_launch(const MegaParams& p, const void* k_pool, const void* v_pool, long long pool_rows, int num_sms,
                       cudaStream_t s);

}  // namespace lcc
