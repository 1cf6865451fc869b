// token-major "augmented head operand" shared by the attention kernels (mirror of gb200_head_operand)
#pragma once
#include "common.cuh"

namespace gb200 {

struct HeadOperand {
    const float* ptr;    // token-major rows
    int ld;              // row stride (floats)
    int col0;            // first column of head 0
    int augmented;       // 1: rows already hold [pos | x] per head (stride d per head)
    const float* gamma;  // (H, d_k) or null
    const float* beta;   // (H, d_k) or null
};

// element i of [ pos[t, 0..p) | gamma[h] * x[t, h, :] + beta[h] ]
__device__ __forceinline__ float load_aug(const HeadOperand& op, const float* __restrict__ pos, int p, int dk, int h,
                                          long long t, int i) {
    if (op.augmented) return op.ptr[t * op.ld + op.col0 + h * (p + dk) + i];
    if (i < p) return pos[t * p + i];
    const int c = i - p;
    float v = op.ptr[t * op.ld + op.col0 + h * dk + c];
    if (op.gamma) v = v * op.gamma[h * dk + c] + op.beta[h * dk + c];
    return v;
}

static inline HeadOperand make_head_operand(const gb200_head_operand* o) {
    HeadOperand h;
    if (!o) { h.ptr = nullptr; h.ld = 0; h.col0 = 0; h.augmented = 0; h.gamma = nullptr; h.beta = nullptr; return h; }
    h.ptr = o->ptr; h.ld = o->ld; h.col0 = o->col0; h.augmented = o->augmented; h.gamma = o->gamma; h.beta = o->beta;
    return h;
}

}  // namespace gb200
