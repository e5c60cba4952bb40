// Internal interface of resize.cu (antialiased bicubic uint8 resize; include/livecc_b200.h lcc_resize_*).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcc {
struct ResizePlan;
int resize_aa_taps(int in_size, int out_size);
// xmin/xsize: int32[out_size]; weights: float[taps][out_size] (tap-major), zero past xsize
int resize_aa_table(int in_size, int out_size, int32_t* xmin, int32_t* xsize, float* weights);
ResizePlan* resize_plan_create(int h, int w, int H, int W, int rows_per_cta /* 0 = choose */);
void resize_plan_destroy(ResizePlan* pl);
void resize_plan_info(const ResizePlan* pl, int* rows_per_cta, int* max_rows, int64_t* smem_bytes);
int resize_bicubic_aa_u8(const ResizePlan* pl, const uint8_t* src, int planes, uint8_t* dst, cudaStream_t s);
}  // namespace lcc
