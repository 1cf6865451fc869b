// Warp-level TF32 tensor-core variants of the attention-core contractions (tf32 precision mode).
//
// The d x d contraction K~^T V~ (d = d_k + pos_dim = 34 at Darcy) is far below any tcgen05 tile
// (M >= 64, TMA-aligned operands) and is HBM/latency-bound, so it runs on mma.sync.m16n8k8 with the
// accumulators in registers, right-sized to d: 3 x 5 tiles for d = 34 instead of the 64 x 64 SIMT tile.
// The "augmented operand" (position columns, per-head LayerNorm affine) is still built on the fly while
// staging tokens into shared memory, rounded to TF32 with cvt.rna.
//
// Shared-memory feature order: the d_k projected features come FIRST (columns 0..d_k-1), the position
// coordinates after them (d_k..d-1), zero padding up to the MMA tile.  That keeps every head row a run of
// 16-byte aligned float4s on both sides of the staging (global rows of Q|K|V are d_k contiguous floats per head)
// and makes the d_k output columns of a token one contiguous run.  perm_col / unperm_col translate to the
// reference's [pos | features] order (libs/layers.py:872-874) wherever a d x d matrix or an augmented row crosses
// the kernel boundary.
#pragma once
#include "common.cuh"
#include "mma_tf32.cuh"

namespace gb200 {

__device__ __forceinline__ int perm_col(int i, int p, int dk) { return i < p ? dk + i : i - p; }     // natural -> smem
__device__ __forceinline__ int unperm_col(int s, int p, int dk) { return s < dk ? p + s : s - dk; }  // smem -> natural

// Stage TT token rows of one head's augmented operand into shared memory S[r][s] (pitch P floats, P % 4 == 0),
// columns s < DPAD (zero beyond d and beyond the last token), rounded to TF32.
//
// Fast path (plain Q|K|V block, d_k % 4 == 0, 16-byte aligned rows): d_k/4 lanes own one token row as float4s, so a
// warp instruction moves 32/(d_k/4) whole rows and gamma/beta are one float4 per thread; the p position columns and
// the zero padding are filled by a second, tiny pass.
// Generic path (rows that already hold [pos | x] per head, i.e. the incoming gradient of the attention output, or
// unaligned views): lane = natural feature column, one scalar per element.
template <int DPAD, int P, int NTHR, int TT>
__device__ __forceinline__ void stage_aug(float* __restrict__ S, const HeadOperand& op, const float* __restrict__ pos,
                                          int p, int dk, int h, long long tok0, int nt) {
    const int d = p + dk;
    const int lq = dk >> 2;                                   // float4s per row
    const bool fast = !op.augmented && (dk & 3) == 0 && lq >= 1 && (NTHR % lq) == 0 && (op.ld & 3) == 0 &&
                      ((op.col0 + h * dk) & 3) == 0 && (reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0 &&
                      (!op.gamma ||
                       ((reinterpret_cast<uintptr_t>(op.gamma) | reinterpret_cast<uintptr_t>(op.beta)) & 15) == 0);
    if (fast) {
        const int rpi = NTHR / lq;                            // rows per pass over the CTA
        const int q = threadIdx.x % lq, r0 = threadIdx.x / lq;
        float4 gm = make_float4(1.f, 1.f, 1.f, 1.f), bt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (op.gamma) {
            gm = *reinterpret_cast<const float4*>(op.gamma + h * dk + 4 * q);
            bt = *reinterpret_cast<const float4*>(op.beta + h * dk + 4 * q);
        }
        const float* src = op.ptr + tok0 * op.ld + op.col0 + h * dk + 4 * q;
        for (int base = 0; base < TT; base += 4 * rpi) {      // four rows per thread in flight
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = base + r0 + u * rpi;
                v[u] = (r < nt) ? *reinterpret_cast<const float4*>(src + (long long)r * op.ld)
                                : make_float4(0.f, 0.f, 0.f, 0.f);      // nt <= TT
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = base + r0 + u * rpi;
                if (r < TT) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nt)
                        o = make_float4(to_tf32(fmaf(v[u].x, gm.x, bt.x)), to_tf32(fmaf(v[u].y, gm.y, bt.y)),
                                        to_tf32(fmaf(v[u].z, gm.z, bt.z)), to_tf32(fmaf(v[u].w, gm.w, bt.w)));
                    *reinterpret_cast<float4*>(&S[r * P + 4 * q]) = o;
                }
            }
        }
        // position columns and zero padding: columns dk .. DPAD-1
        const int extra = DPAD - dk;
        for (int e = threadIdx.x; e < TT * extra; e += NTHR) {
            const int r = e / extra, j = e % extra;
            S[r * P + dk + j] = (j < p && r < nt) ? to_tf32(pos[(tok0 + r) * p + j]) : 0.f;
        }
        return;
    }
    constexpr int NH = (DPAD + 31) / 32;
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const float* src[NH];
    long long pitch[NH];
    float gm[NH], bt[NH];
    int col[NH];
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        const int i = lane + 32 * k;
        src[k] = nullptr; pitch[k] = 0; gm[k] = 1.f; bt[k] = 0.f;
        col[k] = i < d ? perm_col(i, p, dk) : i;
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
            if (i < DPAD) S[r * P + col[k]] = (src[k] && r < nt) ? to_tf32(fmaf(v[it][k], gm[k], bt[k])) : 0.f;
        }
    }
}

// ---- xty: P[i][j] = sum_t L~[t][i] R~[t][j];  MT x NT tiles of 16 x 8, i < 16*MT, j < 8*NT --------------
template <int MT, int NT>
__global__ void __launch_bounds__(256, 4) xty_mma_kernel(HeadOperand L, HeadOperand R, const float* __restrict__ pos,
                                                      int p, int dk, int H, int n, int nsplit, int chunk,
                                                      float* __restrict__ part) {
    pdl_enter();
    constexpr int TC = 64;                         // tokens per stage (8 k-steps of the m16n8k8 MMA)
    constexpr int DM = 16 * MT, DN = 8 * NT;
    // fragment reads index [token = t0 + (lane%4)(+4)][feature = f0 + lane/4]; pitch = 8 (mod 32) makes the
    // 32 lanes hit 32 distinct banks
    constexpr int LP = (DM % 32 == 8) ? DM : DM + ((8 - DM % 32 + 32) % 32);
    constexpr int RP = (DN % 32 == 8) ? DN : DN + ((8 - DN % 32 + 32) % 32);
    __shared__ __align__(16) float Ls[TC][LP];
    __shared__ __align__(16) float Rs[TC][RP];
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
            const int si = 16 * a + g + (k >= 2 ? 8 : 0), sj = 8 * c + 2 * tq + (k & 1);
            if (si < d && sj < d) out[unperm_col(si, p, dk) * d + unperm_col(sj, p, dk)] = acc[sl][k];
        }
    }
}

// ---- xm: O[t][j] = sum_i L~[t][i] M[i][j]  (or M^T);  CTA = TT tokens (TT/16 warps x 16), KT x NT tiles -------
template <int KT, int NT, int TT>
__global__ void __launch_bounds__(TT * 2, 1024 / (TT * 2)) xm_mma_kernel(HeadOperand L, const float* __restrict__ pos,
                                                     const float* __restrict__ Mat, int transM, int p, int dk, int H,
                                                     int n, float* __restrict__ out, int ldo, int ocol0,
                                                     int out_augmented, float oscale) {
    pdl_enter();
    constexpr int DK = 8 * KT, DN = 8 * NT, NTHR = TT * 2;   // TT/16 warps, 16 tokens each
    constexpr int LP = (DK % 32 == 4) ? DK : DK + ((4 - DK % 32 + 32) % 32);    // A frag: [token g][feat tq]
    constexpr int MP = (DN % 32 == 8) ? DN : DN + ((8 - DN % 32 + 32) % 32);    // B frag: [feat tq][col g]
    __shared__ __align__(16) float Ls[TT][LP];
    __shared__ __align__(16) float Ms[DK][MP];
    const int d = p + dk;
    const int bh = blockIdx.y, b = bh / H, h = bh % H;
    const int t0 = blockIdx.x * TT;
    const int nt = min(TT, n - t0);
    const float* M = Mat + (long long)bh * d * d;
    // the d x d matrix in the shared-memory feature order on both axes (zero padded); loads batched ahead of stores
    {
        constexpr int PER = (DK * DN + NTHR - 1) / NTHR;
        float mv[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int e = threadIdx.x + u * NTHR;
            const int si = e / DN, sj = e % DN;
            mv[u] = 0.f;
            if (e < DK * DN && si < d && sj < d) {
                const int i = unperm_col(si, p, dk), j = unperm_col(sj, p, dk);
                mv[u] = transM ? M[j * d + i] : M[i * d + j];
            }
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int e = threadIdx.x + u * NTHR;
            if (e < DK * DN) Ms[e / DN][e % DN] = to_tf32(mv[u]);
        }
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
    // Each thread owns column pairs (sj, sj + 1), sj even.  Projected-feature pairs (sj + 1 < d_k) are adjacent in both
    // output layouts, so they leave as one 8-byte store when the row base is 8-byte aligned (4 lanes = one 32-byte
    // sector per row); position columns (augmented outputs only) are scalar.
    const int fbase = ocol0 + (out_augmented ? h * d + p : h * dk);          // output column of feature 0
    const bool pair_ok = ((fbase | ldo) & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
#pragma unroll
    for (int c = 0; c < NT; ++c)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = tr + g + 8 * half;
            const int sj = 8 * c + 2 * tq;
            if (r >= nt || sj >= d) continue;
            const long long t = (long long)b * n + t0 + r;
            const float v0 = acc[c][2 * half] * oscale, v1 = acc[c][2 * half + 1] * oscale;
            if (sj + 1 < dk && pair_ok) {
                *reinterpret_cast<float2*>(out + t * ldo + fbase + sj) = make_float2(v0, v1);
            } else {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int s = sj + u;
                    if (s >= d) continue;
                    const float v = u ? v1 : v0;
                    if (s < dk) out[t * ldo + fbase + s] = v;
                    else if (out_augmented) out[t * ldo + ocol0 + h * d + (s - dk)] = v;
                }
            }
        }
}

}  // namespace gb200
