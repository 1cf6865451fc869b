// Warp-level TF32 tensor-core variants of the attention-core contractions (tf32 precision mode).
//
// The d x d contraction K~^T V~ (d = d_k + pos_dim = 34 at Darcy) is far below any tcgen05 tile
// (M >= 64, TMA-aligned operands) and is HBM/latency-bound, so it runs on mma.sync.m16n8k8 with the
// accumulators in registers, right-sized to d: 3 x 5 tiles for d = 34 instead of the 64 x 64 SIMT tile.
// The "augmented operand" (position columns, per-head LayerNorm affine) is still built on the fly while
// staging tokens into shared memory, rounded to TF32 with cvt.rna.
#pragma once
#include "common.cuh"
#include "mma_tf32.cuh"

namespace gb200 {

// Stage TT token rows of one head's augmented operand into shared memory S[r][i] (pitch P floats), features
// i < DPAD (zero beyond d and beyond the last token), rounded to TF32.  Lane = feature column, so the column
// kind (position / normalised feature / padding), its affine (gamma, beta) and its source pointer are resolved
// once per thread; the token loop is then one coalesced load + FMA + convert + store per element.
template <int DPAD, int P, int NTHR, int TT>
__device__ __forceinline__ void stage_aug(float* __restrict__ S, const HeadOperand& op, const float* __restrict__ pos,
                                          int p, int dk, int h, long long tok0, int nt) {
    constexpr int NH = (DPAD + 31) / 32;
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int d = p + dk;
    const float* src[NH];
    long long pitch[NH];
    float gm[NH], bt[NH];
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        const int i = lane + 32 * k;
        src[k] = nullptr; pitch[k] = 0; gm[k] = 1.f; bt[k] = 0.f;
        if (i < d) {
            if (op.augmented) { src[k] = op.ptr + tok0 * op.ld + op.col0 + h * d + i; pitch[k] = op.ld; }
            else if (i < p) { src[k] = pos + tok0 * p + i; pitch[k] = p; }
            else {
                const int c = i - p;
                src[k] = op.ptr + tok0 * op.ld + op.col0 + h * dk + c; pitch[k] = op.ld;
                if (op.gamma) { gm[k] = op.gamma[h * dk + c]; bt[k] = op.beta[h * dk + c]; }
            }
        }
    }
    // all of this warp's loads are issued before the first store (ITER x NH independent requests in flight)
    constexpr int NW = NTHR / 32, ITER = TT / NW;
    static_assert(TT % NW == 0, "tokens per stage must divide over the warps");
    float v[ITER][NH];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int r = warp + it * NW;
#pragma unroll
        for (int k = 0; k < NH; ++k) v[it][k] = (src[k] && r < nt) ? src[k][r * pitch[k]] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int r = warp + it * NW;
#pragma unroll
        for (int k = 0; k < NH; ++k) {
            const int i = lane + 32 * k;
            if (i < DPAD) S[r * P + i] = (src[k] && r < nt) ? to_tf32(fmaf(v[it][k], gm[k], bt[k])) : 0.f;
        }
    }
}

// ---- xty: P[i][j] = sum_t L~[t][i] R~[t][j];  MT x NT tiles of 16 x 8, i < 16*MT, j < 8*NT --------------
template <int MT, int NT>
__global__ void __launch_bounds__(256, 3) xty_mma_kernel(HeadOperand L, HeadOperand R, const float* __restrict__ pos,
                                                      int p, int dk, int H, int n, int nsplit, int chunk,
                                                      float* __restrict__ part) {
    pdl_enter();
    constexpr int TC = 64;                         // tokens per stage (8 k-steps of the m16n8k8 MMA)
    constexpr int DM = 16 * MT, DN = 8 * NT;
    constexpr int LS = DM + 8 - (DM % 32 == 8 ? 0 : 0), RS = DN;   // pitches: see bank note below
    // fragment reads index [token = t0 + (lane%4)(+4)][feature = f0 + lane/4]; pitch = 8 (mod 32) makes the
    // 32 lanes hit 32 distinct banks
    constexpr int LP = (DM % 32 == 8) ? DM : DM + ((8 - DM % 32 + 32) % 32);
    constexpr int RP = (DN % 32 == 8) ? DN : DN + ((8 - DN % 32 + 32) % 32);
    __shared__ float Ls[TC][LP];
    __shared__ float Rs[TC][RP];
    (void)LS; (void)RS;
    const int d = p + dk;
    const int bh = blockIdx.y, b = bh / H, h = bh % H, split = blockIdx.x;
    const int tbeg = split * chunk, tend = min(n, tbeg + chunk);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, tq = lane % 4;
    // Warps split the OUTPUT tiles (tile = warp, warp + 8, ...) and each contracts all staged tokens: no
    // cross-warp reduction, a handful of accumulator registers per thread.
    constexpr int TILES = MT * NT, SLOTS = (TILES + 7) / 8;
    float acc[SLOTS][4];
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[sl][k] = 0.f;

    for (int t0 = tbeg; t0 < tend; t0 += TC) {
        const int nt = min(TC, tend - t0);
        stage_aug<DM, LP, 256, TC>(&Ls[0][0], L, pos, p, dk, h, (long long)b * n + t0, nt);
        stage_aug<DN, RP, 256, TC>(&Rs[0][0], R, pos, p, dk, h, (long long)b * n + t0, nt);
        __syncthreads();
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
            const int tile = warp + 8 * sl;
            if (tile < TILES) {
                const int a = tile / NT, c = tile % NT;
#pragma unroll
                for (int ks = 0; ks < TC / 8; ++ks) {
                    const int tk = 8 * ks;
                    float af[4], bf[2];
                    af[0] = Ls[tk + tq][16 * a + g];
                    af[1] = Ls[tk + tq][16 * a + g + 8];
                    af[2] = Ls[tk + tq + 4][16 * a + g];
                    af[3] = Ls[tk + tq + 4][16 * a + g + 8];
                    bf[0] = Rs[tk + tq][8 * c + g];
                    bf[1] = Rs[tk + tq + 4][8 * c + g];
                    mma_tf32(acc[sl], af, bf);
                }
            }
        }
        __syncthreads();
    }
    float* out = part + ((long long)bh * nsplit + split) * d * d;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
        const int tile = warp + 8 * sl;
        if (tile >= TILES) continue;
        const int a = tile / NT, c = tile % NT;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 16 * a + g + (k >= 2 ? 8 : 0), j = 8 * c + 2 * tq + (k & 1);
            if (i < d && j < d) out[i * d + j] = acc[sl][k];
        }
    }
}

// ---- xm: O[t][j] = sum_i L~[t][i] M[i][j]  (or M^T);  CTA = TT tokens (TT/16 warps x 16), KT x NT tiles -------
template <int KT, int NT, int TT>
__global__ void __launch_bounds__(TT * 2) xm_mma_kernel(HeadOperand L, const float* __restrict__ pos,
                                                     const float* __restrict__ Mat, int transM, int p, int dk, int H,
                                                     int n, float* __restrict__ out, int ldo, int ocol0,
                                                     int out_augmented, float oscale) {
    pdl_enter();
    constexpr int DK = 8 * KT, DN = 8 * NT, NTHR = TT * 2;   // TT/16 warps, 16 tokens each
    constexpr int LP = (DK % 32 == 4) ? DK : DK + ((4 - DK % 32 + 32) % 32);    // A frag: [token g][feat tq]
    constexpr int MP = (DN % 32 == 8) ? DN : DN + ((8 - DN % 32 + 32) % 32);    // B frag: [feat tq][col g]
    __shared__ float Ls[TT][LP];
    __shared__ float Ms[DK][MP];
    const int d = p + dk;
    const int bh = blockIdx.y, b = bh / H, h = bh % H;
    const int t0 = blockIdx.x * TT;
    const int nt = min(TT, n - t0);
    const float* M = Mat + (long long)bh * d * d;
    for (int e = threadIdx.x; e < DK * DN; e += NTHR) {
        const int i = e / DN, j = e % DN;
        float v = 0.f;
        if (i < d && j < d) v = to_tf32(transM ? M[j * d + i] : M[i * d + j]);
        Ms[i][j] = v;
    }
    stage_aug<DK, LP, NTHR, TT>(&Ls[0][0], L, pos, p, dk, h, (long long)b * n + t0, nt);
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, tq = lane % 4;
    const int tr = warp * 16;
    float acc[NT][4];
#pragma unroll
    for (int c = 0; c < NT; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[c][k] = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        float af[4];
        af[0] = Ls[tr + g][8 * k + tq];
        af[1] = Ls[tr + g + 8][8 * k + tq];
        af[2] = Ls[tr + g][8 * k + tq + 4];
        af[3] = Ls[tr + g + 8][8 * k + tq + 4];
#pragma unroll
        for (int c = 0; c < NT; ++c) {
            float bf[2];
            bf[0] = Ms[8 * k + tq][8 * c + g];
            bf[1] = Ms[8 * k + tq + 4][8 * c + g];
            mma_tf32(acc[c], af, bf);
        }
    }
#pragma unroll
    for (int c = 0; c < NT; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = tr + g + (k >= 2 ? 8 : 0);
            const int j = 8 * c + 2 * tq + (k & 1);
            if (r >= nt || j >= d) continue;
            const long long t = (long long)b * n + t0 + r;
            const float v = acc[c][k] * oscale;
            if (out_augmented) out[t * ldo + ocol0 + h * d + j] = v;
            else if (j >= p) out[t * ldo + ocol0 + h * dk + (j - p)] = v;
        }
}

}  // namespace gb200
