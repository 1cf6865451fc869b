// Fused backward of the Galerkin encoder layer (the adjoint of csrc/encoder_fwd.cu; reference forward:
// libs/model.py:104-140, libs/layers.py:829-899, 979-987).  CTA = one 128-token tile of one sample; four tcgen05 kernels
// carry every input-gradient GEMM (bf16x3, encoder_common.cuh), the elementwise backward passes and the bias / LayerNorm
// parameter reductions; the four weight-gradient GEMMs (contractions over ALL tokens) stay on the library's TF32 GEMM.
//
//   enc_ffn_bwd_kernel   g2 = dy * mask2 ; g1 = (g2 W2) * [hidden > 0] / (1 - p_ffn) ; dx1 = dy + g1 W1
//                        + column sums of g2, g1 (bias gradients) as tensor-core "ones" contractions over the tile
//   enc_attn_bwd_kernel  g_fc = sign * dx1 * mask1 ; dheads = g_fc W_fc ; dQ = dheads_h A_h^T ;
//                        partial G = Q~^T dheads (adjoint of the K~^T V~ contraction's consumer) ; column sums of g_fc
//   enc_kv_bwd_kernel    dA = mask * scale * sum(G partials) ; dV~ = K~ dA ; dK~ = V~ dA^T ; per-head LayerNorm backward
//                        -> dqkv[:, K | V] ; partial d gamma / d beta
//   enc_dx_kernel        dx = dqkv W_qkv + dx1 ; column sums of dqkv
//   enc_reduce_kernel    fixed-order sum of the per-tile partial vectors (deterministic)
#include <cuda.h>

#include "encoder_common.cuh"

namespace gb200 {
namespace enc {

// ---- shared device helpers ------------------------------------------------------------------------------------------
// landed fp32 tile (four SWIZZLE_128B blocks): v = scale * dropmask(p) * v in place, optionally also to global (coalesced)
__device__ __forceinline__ void mask_tile(uint8_t* tile, int wt, float p, unsigned long long seed, long long grow0,
                                          int nvalid, float scale, float* gout) {
    for (int idx = wt; idx < TM * 32; idx += NWORK * 32) {
        const int row = idx >> 5, c4 = idx & 31;
        float4* ptr = reinterpret_cast<float4*>(tile + (c4 >> 3) * TILE_BYTES + unit_off(row, c4 & 7));
        float4 v = *ptr;
        if (p > 0.f) {
            const float4 ds = dropout_scale4(p, seed, (unsigned long long)(grow0 + row) * DM + c4 * 4);
            v.x *= ds.x; v.y *= ds.y; v.z *= ds.z; v.w *= ds.w;
        }
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        *ptr = v;
        if (gout && row < nvalid) *reinterpret_cast<float4*>(gout + (grow0 + row) * DM + c4 * 4) = v;
    }
}
// [128 token rows][16 columns of 1.0] as an MN-major bf16 block (units 0, 1 of every row)
__device__ __forceinline__ void write_ones(uint8_t* blk, int wt) {
    if (wt < TM) {
        const uint4 one = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        *reinterpret_cast<uint4*>(blk + unit_off(wt, 0)) = one;
        *reinterpret_cast<uint4*>(blk + unit_off(wt, 1)) = one;
    }
}
// D[f][0..15] = sum over the tile's tokens of A[token][f]  (A: MN-major hi/lo blocks of 64 features, `lbo` apart)
__device__ __forceinline__ void mma_colsum(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t lbo, uint32_t ones) {
    const uint32_t id16 = idesc_bf16_mn(16);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        tc_mma_bf16(d, mndesc(a_hi + ks * 2048, lbo), mndesc(ones + ks * 2048, lbo), id16, ks == 0 ? 0u : 1u);
        tc_mma_bf16(d, mndesc(a_lo + ks * 2048, lbo), mndesc(ones + ks * 2048, lbo), id16, 1u);
    }
}
// as split_tile_inplace, with v = v * gamma[col] + beta[col] applied first (gamma == null: plain split)
__device__ __forceinline__ void split_tile_affine(uint8_t* tile, int r, int kc, const float* gamma, const float* beta) {
    uint8_t* ba = tile + (2 * kc) * TILE_BYTES;
    uint8_t* bb = ba + TILE_BYTES;
    float4 v[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        v[c] = *reinterpret_cast<const float4*>(ba + unit_off(r, c));
        v[8 + c] = *reinterpret_cast<const float4*>(bb + unit_off(r, c));
    }
    if (gamma) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + 64 * kc + 4 * c);
            const float4 b = *reinterpret_cast<const float4*>(beta + 64 * kc + 4 * c);
            v[c].x = fmaf(v[c].x, g.x, b.x); v[c].y = fmaf(v[c].y, g.y, b.y);
            v[c].z = fmaf(v[c].z, g.z, b.z); v[c].w = fmaf(v[c].w, g.w, b.w);
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        uint4 h, l;
        split2(v[2 * u].x, v[2 * u].y, h.x, l.x);
        split2(v[2 * u].z, v[2 * u].w, h.y, l.y);
        split2(v[2 * u + 1].x, v[2 * u + 1].y, h.z, l.z);
        split2(v[2 * u + 1].z, v[2 * u + 1].w, h.w, l.w);
        *reinterpret_cast<uint4*>(ba + unit_off(r, u)) = h;
        *reinterpret_cast<uint4*>(bb + unit_off(r, u)) = l;
    }
}
// ---------------------------------------------------------------------------------------------------------------
// kernel B1: feed-forward backward
// ---------------------------------------------------------------------------------------------------------------
struct FfnBwdArgs {
    const uint8_t* w2t;         // 8 tiles of W2^T: (N block 0/1 of the hidden index, K chunk 0/1 of the output index, hi/lo)
    const uint8_t* w1t;         // 8 tiles of W1^T: (K chunk 0..3 of the hidden index, hi/lo)
    const float* dy;            // (B n, 128)
    const float* hid;           // (B n, 256) saved drop(relu(.))
    float* g2;                  // (B n, 128) = dy * mask2 (null when p2 == 0: dy itself is g2)
    float* g1;                  // (B n, 256)
    float* dx1;                 // (B n, 128)
    float* part;                // (B tiles, VEC_FLOATS) per-tile column sums
    float p2, pf;
    unsigned long long seed2;
    const unsigned long long* seed_off;
    int B, n, tiles;
};
constexpr int FB_RING = 2;
constexpr int FB_SMEM = 8 * TILE_BYTES + FB_RING * TILE_BYTES + TILE_BYTES + NWORK * STAGE_BYTES + 512 + 1024;

__global__ void __launch_bounds__(THREADS, 1) enc_ffn_bwd_kernel(const __grid_constant__ CUtensorMap mapDY, FfnBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_base(smem_raw);
    uint8_t* R = smem;                                         // dy tile (4 tiles) -> g1 operand (8 tiles)
    uint8_t* ring = R + 8 * TILE_BYTES;
    uint8_t* ones = ring + FB_RING * TILE_BYTES;
    float* staging = reinterpret_cast<float*>(ones + TILE_BYTES);
    Bars* bar = reinterpret_cast<Bars*>(staging + NWORK * 32 * STAGE_PITCH);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x % a.tiles;
    const int t0 = tile * TM;
    const int nvalid = min(TM, a.n - t0);

    if (threadIdx.x == 0) {
        mbar_init(&bar->xfull, 1);
        mbar_init(&bar->xconv, NWORK);
        for (int s = 0; s < FB_RING; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar->dfull[i], 1); mbar_init(&bar->hand[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapDY) : "memory");
    }
    if (threadIdx.x == 0) { trace_entry(); trace(0); }
    const uint32_t tmem = tmem_alloc_512(&bar->tmem_slot, warp);
    grid_dep_sync();
    constexpr int DX_COL = 256, DB2_COL = 384, DB1A_COL = 400, DB1B_COL = 416;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&bar->xfull, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(R + j * TILE_BYTES, &mapDY, &bar->xfull, 32 * j, t0, b);
            for (int i = 0; i < 16; ++i) {
                const int s = i % FB_RING;
                mbar_wait(&bar->empty[s], ((i / FB_RING) & 1) ^ 1);
                mbar_expect_tx(&bar->full[s], TILE_BYTES);
                const uint8_t* src = i < 8 ? a.w2t + (size_t)i * TILE_BYTES : a.w1t + (size_t)(i - 8) * TILE_BYTES;
                bulk_load(ring + s * TILE_BYTES, src, TILE_BYTES, &bar->full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ra = smem_u32(R), rb = smem_u32(ring), on = smem_u32(ones);
            const uint32_t id128 = idesc_bf16(128);
            mbar_wait(&bar->xconv, 0);
            tc_fence_after();
            trace(4);
            mma_colsum(tmem + DB2_COL, ra, ra + TILE_BYTES, 2 * TILE_BYTES, on);          // d b2 = colsum(g2)
            for (int i = 0; i < 8; ++i) {                      // g1_pre = g2 W2
                const int s = i % FB_RING;
                mbar_wait(&bar->full[s], (i / FB_RING) & 1);
                tc_fence_after();
                const int nb = i >> 2, kc = (i >> 1) & 1;
                mma_weight_tile(tmem + nb * 128, ra + (2 * kc) * TILE_BYTES, ra + (2 * kc + 1) * TILE_BYTES,
                                rb + s * TILE_BYTES, (i & 1) == 0, id128, kc == 0);
                tc_commit(&bar->empty[s]);
                if ((i & 3) == 3) tc_commit(&bar->dfull[nb]);
            }
            trace(5);
            for (int i = 8; i < 16; ++i) {                     // dx1_pre = g1 W1
                const int s = i % FB_RING;
                const int kc = (i - 8) >> 1;
                if (((i - 8) & 1) == 0) {
                    mbar_wait(&bar->hand[kc], 0);
                    tc_fence_after();
                    // d b1 = colsum(g1): features 0..127 once chunks 0,1 are in place, 128..255 after chunks 2,3
                    if (kc == 1) mma_colsum(tmem + DB1A_COL, ra, ra + TILE_BYTES, 2 * TILE_BYTES, on);
                    if (kc == 3) mma_colsum(tmem + DB1B_COL, ra + 4 * TILE_BYTES, ra + 5 * TILE_BYTES, 2 * TILE_BYTES, on);
                }
                mbar_wait(&bar->full[s], (i / FB_RING) & 1);
                tc_fence_after();
                mma_weight_tile(tmem + DX_COL, ra + (2 * kc) * TILE_BYTES, ra + (2 * kc + 1) * TILE_BYTES,
                                rb + s * TILE_BYTES, (i & 1) == 0, id128, kc == 0);
                tc_commit(&bar->empty[s]);
            }
            tc_commit(&bar->dfull[2]);
            trace(8);
        }
    } else {
        const int w = warp - 2, q = warp & 3, hf = w >> 2;
        const int wt = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const long long grow0t = (long long)b * a.n + t0;
        const long long grow0 = grow0t + q * 32;
        const int nrows = max(0, min(32, nvalid - q * 32));
        float* stage = staging + w * 32 * STAGE_PITCH;
        const unsigned long long so = a.seed_off ? *a.seed_off : 0ull;
        write_ones(ones, wt);
        mbar_wait(&bar->xfull, 0);
        if (wt == 0) trace(9);
        if (a.p2 > 0.f) {
            mask_tile(R, wt, a.p2, a.seed2 + so, grow0t, nvalid, 1.f, a.g2);
            worker_bar();
        }
        split_tile_inplace(R, wt & 127, wt >> 7);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->xconv);
        if (wt == 0) trace(10);

        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const float keepf = a.pf > 0.f ? dropout_keep_scale(dropout_threshold(a.pf)) : 1.f;
        float v[32];
        mbar_wait(&bar->dfull[hf], 0);
        tc_fence_after();
        if (wt == 0) trace(11);
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = hf * 128 + cc * 32;
            float r[32];
            warp_load_block(stage, r, lane, a.hid + grow0 * DFF + c0, DFF, nrows);
            tmem_ld32(tlane + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = r[j] > 0.f ? v[j] * keepf : 0.f;
            warp_store_block(stage, v, lane, a.g1 + grow0 * DFF + c0, DFF, nrows);
            if (cc == 0 && hf == 0) {                          // chunks 0,1 alias the g2 operand
                mbar_wait(&bar->dfull[1], 0);
                tc_fence_after();
            }
            const int kc = c0 >> 6;
            uint8_t* hi = R + (2 * kc) * TILE_BYTES;
#pragma unroll
            for (int u = 0; u < 4; ++u) store_unit(hi, hi + TILE_BYTES, row, ((c0 & 63) >> 3) + u, &v[8 * u]);
            if (cc & 1) {
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar->hand[kc]);
            }
        }
        if (wt == 0) trace(13);
        mbar_wait(&bar->dfull[2], 0);
        tc_fence_after();
        if (wt == 0) trace(15);
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = hf * 64 + cc * 32;
            float r[32];
            warp_load_block(stage, r, lane, a.dy + grow0 * DM + c0, DM, nrows);
            tmem_ld32(tlane + DX_COL + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += r[j];
            warp_store_block(stage, v, lane, a.dx1 + grow0 * DM + c0, DM, nrows);
        }
        // bias-gradient partials of this tile (lane = feature)
        float* P = a.part + (long long)(b * a.tiles + tile) * VEC_FLOATS;
        float s0, s1;
        if (hf == 0) {
            tmem_ld2(tlane + DB2_COL, s0, s1);
            P[VEC_B2 + row] = s0;
        } else {
            tmem_ld2(tlane + DB1A_COL, s0, s1);
            P[VEC_B1 + row] = s0;
            tmem_ld2(tlane + DB1B_COL, s0, s1);
            P[VEC_B1 + 128 + row] = s0;
        }
        tc_fence_before();
        if (wt == 0) trace(16);
    }
    tmem_free_512(tmem, warp);
    if (threadIdx.x == 0) trace(17);
}

// ---------------------------------------------------------------------------------------------------------------
// kernel B2: fc backward, dQ, partial G
// ---------------------------------------------------------------------------------------------------------------
struct AttnBwdArgs {
    const uint8_t* fct;         // 8 tiles of W_fc'^T: (N block 0/1 of the head-padded index, K chunk 0/1 of the output, hi/lo)
    const float* dx1;           // (B n, 128)
    float* gfc;                 // (B n, 128) = sign * dx1 * mask1, or null when it equals dx1
    float p1, sign;
    unsigned long long seed1;
    const unsigned long long* seed_off;
    const float* attn;          // (B, 4, d, d) masked, scaled attention matrix of the forward
    const float* pos;
    float* dqkv;                // (B n, 384): columns 0..127 written here
    float* gpart;               // (B, tiles, 4, d, d)
    float* part;
    int B, n, p, tiles;
};
constexpr int AB_RING = 2;
constexpr int AB_SMEM = 6 * TILE_BYTES + 4 * TILE_BYTES + AB_RING * TILE_BYTES + 2 * TILE_BYTES + 512 + 1024;

__global__ void __launch_bounds__(THREADS, 1) enc_attn_bwd_kernel(const __grid_constant__ CUtensorMap mapDX,
                                                                  const __grid_constant__ CUtensorMap mapQ, AttnBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_base(smem_raw);
    uint8_t* AO = smem;                                        // dx1 tile (4) | ones (1) | spare  ->  dheads operand (6)
    uint8_t* ones = AO + 4 * TILE_BYTES;
    uint8_t* R2 = AO + 6 * TILE_BYTES;                         // Q tile
    uint8_t* ring = R2 + 4 * TILE_BYTES;                       // W_fc^T tiles, later the A_h^T operand of dQ
    uint8_t* posb = ring + AB_RING * TILE_BYTES;               // [tokens][pos] MN-major block, hi | lo
    Bars* bar = reinterpret_cast<Bars*>(posb + 2 * TILE_BYTES);
    float* staging = reinterpret_cast<float*>(AO);             // final epilogue only (every operand is dead by then)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x % a.tiles;
    const int t0 = tile * TM;
    const int nvalid = min(TM, a.n - t0);
    const int p = a.p, d = DK + p, dd = d * d;

    if (threadIdx.x == 0) {
        mbar_init(&bar->xfull, 1);
        mbar_init(&bar->xconv, NWORK);
        mbar_init(&bar->done, 1);                              // Q tile landed
        for (int s = 0; s < AB_RING; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar->dfull[i], 1); mbar_init(&bar->hand[i], NWORK); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapDX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQ) : "memory");
    }
    if (threadIdx.x == 0) { trace_entry(); trace(0); }
    const uint32_t tmem = tmem_alloc_512(&bar->tmem_slot, warp);
    grid_dep_sync();
    constexpr int DQ_COL = 192, DBFC_COL = 320, GP1_COL = 336, GP2_COL = 352;
    constexpr uint32_t BQH = 8192;                             // per-head A_h^T image: hi 4 KB | lo 4 KB

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&bar->xfull, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(AO + j * TILE_BYTES, &mapDX, &bar->xfull, 32 * j, t0, b);
            mbar_expect_tx(&bar->done, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(R2 + j * TILE_BYTES, &mapQ, &bar->done, 32 * j, t0, b);
            for (int i = 0; i < 8; ++i) {
                const int s = i % AB_RING;
                mbar_wait(&bar->empty[s], ((i / AB_RING) & 1) ^ 1);
                mbar_expect_tx(&bar->full[s], TILE_BYTES);
                bulk_load(ring + s * TILE_BYTES, a.fct + (size_t)i * TILE_BYTES, TILE_BYTES, &bar->full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ao = smem_u32(AO), r2 = smem_u32(R2), rb = smem_u32(ring), on = smem_u32(ones),
                           pb = smem_u32(posb);
            mbar_wait(&bar->xconv, 0);
            tc_fence_after();
            trace(4);
            mma_colsum(tmem + DBFC_COL, ao, ao + TILE_BYTES, 2 * TILE_BYTES, on);          // d b_fc = colsum(g_fc)
            for (int i = 0; i < 8; ++i) {                      // dheads = g_fc W_fc' : 128 + 64 head-padded columns
                const int s = i % AB_RING;
                mbar_wait(&bar->full[s], (i / AB_RING) & 1);
                tc_fence_after();
                const int nb = i >> 2, kc = (i >> 1) & 1;
                mma_weight_tile(tmem + nb * 128, ao + (2 * kc) * TILE_BYTES, ao + (2 * kc + 1) * TILE_BYTES,
                                rb + s * TILE_BYTES, (i & 1) == 0, idesc_bf16(nb == 0 ? 128 : 64), kc == 0);
                tc_commit(&bar->empty[s]);
            }
            tc_commit(&bar->dfull[0]);
            mbar_wait(&bar->hand[0], 0);
            tc_fence_after();
            trace(7);
            const uint32_t id32 = idesc_bf16(32);
#pragma unroll
            for (int h = 0; h < 4; ++h)                        // dQ_h = dheads_h (A_h[p:, :])^T : K = 48 = three k-steps
#pragma unroll
                for (int s3 = 0; s3 < 3; ++s3) {
                    const int k0 = HP * h + 16 * s3;
                    const uint32_t ah = ao + (2 * (k0 >> 6)) * TILE_BYTES + (k0 & 63) * 2, al = ah + TILE_BYTES;
                    const uint32_t bh = rb + h * BQH + s3 * 32, bl = bh + 4096;
                    mma3(tmem + DQ_COL + 32 * h, ah, al, bh, bl, id32, s3 == 0);
                }
            // G = Q^T dheads  (M = 128 Q features, N = 192), Gp = dheads^T pos  (M = 192 as 128 + 64.., N = 16)
            const uint32_t id192 = idesc_bf16_mn(192), id16 = idesc_bf16_mn(16);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t qh = r2 + ks * 2048, ql = qh + TILE_BYTES;
                const uint32_t dh = ao + ks * 2048, dl = dh + TILE_BYTES;
                mma3_mn(tmem, qh, ql, 2 * TILE_BYTES, dh, dl, 2 * TILE_BYTES, id192, ks == 0);
                mma3_mn(tmem + GP1_COL, dh, dl, 2 * TILE_BYTES, pb + ks * 2048, pb + TILE_BYTES + ks * 2048, 2 * TILE_BYTES,
                        id16, ks == 0);
                mma3_mn(tmem + GP2_COL, dh + 4 * TILE_BYTES, dl + 4 * TILE_BYTES, 2 * TILE_BYTES, pb + ks * 2048,
                        pb + TILE_BYTES + ks * 2048, 2 * TILE_BYTES, id16, ks == 0);
            }
            tc_commit(&bar->dfull[1]);
            trace(8);
        }
    } else {
        const int w = warp - 2, q = warp & 3, hf = w >> 2;
        const int wt = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const bool valid = row < nvalid;
        const long long grow0t = (long long)b * a.n + t0;
        const long long grow = grow0t + row;
        const long long grow0 = grow0t + q * 32;
        const int nrows = max(0, min(32, nvalid - q * 32));
        const unsigned long long so = a.seed_off ? *a.seed_off : 0ull;
        float pv[2] = {0.f, 0.f};
        if (valid && hf == 0) {
            if (p > 0) pv[0] = a.pos[grow * p];
            if (p > 1) pv[1] = a.pos[grow * p + 1];
        }
        write_ones(ones, wt);
        mbar_wait(&bar->xfull, 0);
        if (wt == 0) trace(9);
        if (a.gfc) {
            mask_tile(AO, wt, a.p1, a.seed1 + so, grow0t, nvalid, a.sign, a.gfc);
            worker_bar();
        }
        split_tile_inplace(AO, wt & 127, wt >> 7);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->xconv);
        if (wt == 0) trace(10);
        // Q tile: MN-major operand of the G contraction; position block for Gp
        mbar_wait(&bar->done, 0);
        split_tile_inplace(R2, wt & 127, wt >> 7);
        if (hf == 0) {
            float x8[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x8[c] = 0.f;
            x8[0] = pv[0]; x8[1] = pv[1];
            store_unit(posb, posb + TILE_BYTES, row, 0, &x8[0]);
            store_unit(posb, posb + TILE_BYTES, row, 1, &x8[8]);
        }
        if (wt == 0) trace(11);

        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        float v[32];
        mbar_wait(&bar->dfull[0], 0);                          // dheads accumulators ready; dx1 tile, ones, ring are dead
        tc_fence_after();
        if (wt == 0) trace(14);
        // A_h^T operand of dQ: row = Q feature i, K = head column j (48, zero past d), one 4 KB image per head
        {
            const float* Ab = a.attn + (long long)b * NH * dd;
            for (int u = wt; u < 4 * 32 * 6; u += NWORK * 32) {
                const int h = u / 192, i = (u / 6) & 31, uu = u % 6;
                float x8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int j = 8 * uu + k;
                    x8[k] = j < d ? __ldg(Ab + h * dd + (p + i) * d + j) : 0.f;
                }
                store_unit(ring + h * BQH, ring + h * BQH + 4096, i, uu, x8);
            }
        }
        for (int hh = 0; hh < 2; ++hh) {
            const int h = 2 * hf + hh;
            float e0, e1;
            tmem_ld32(tlane + h * HP, v);
            tmem_ld2(tlane + h * HP + 32, e0, e1);
            if (d < 34) e1 = 0.f;
            if (d < 33) e0 = 0.f;
            float tail[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) tail[j] = 0.f;
            tail[0] = e0; tail[1] = e1;
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int g = 6 * h + u;
                uint8_t* hi = AO + (2 * (g >> 3)) * TILE_BYTES;
                store_unit(hi, hi + TILE_BYTES, row, g & 7, u < 4 ? &v[8 * u] : &tail[8 * (u - 4)]);
            }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->hand[0]);
        if (wt == 0) trace(15);

        mbar_wait(&bar->dfull[1], 0);
        tc_fence_after();
        if (wt == 0) trace(16);
        worker_bar();                                          // every warp is past its operand reads: staging may alias AO
        float* stage = staging + w * 32 * STAGE_PITCH;
        for (int hh = 0; hh < 2; ++hh) {
            const int h = 2 * hf + hh;
            tmem_ld32(tlane + DQ_COL + 32 * h, v);
            warp_store_block(stage, v, lane, a.dqkv + grow0 * 384 + 32 * h, 384, nrows);
        }
        float* GP = a.gpart + (long long)(b * a.tiles + tile) * NH * dd;
        float* P = a.part + (long long)(b * a.tiles + tile) * VEC_FLOATS;
        if (hf == 0) {     // rows = Q features of head q: G_q[p + lane][0 .. d)
            float e0, e1;
            tmem_ld32(tlane + HP * q, v);
            tmem_ld2(tlane + HP * q + 32, e0, e1);
            float* prow = GP + q * dd + (p + lane) * d;
#pragma unroll
            for (int j = 0; j < 32; ++j) prow[j] = v[j];
            if (d > 32) prow[32] = e0;
            if (d > 33) prow[33] = e1;
            tmem_ld2(tlane + DBFC_COL, e0, e1);
            P[VEC_BFC + row] = e0;
        } else {           // rows = head-padded dheads columns r = 48 h + j:  Gp[r][c] = G_h[c][j]
            float c0, c1;
            tmem_ld2(tlane + GP1_COL, c0, c1);
            {
                const int h = row / HP, j = row % HP;
                if (j < d) {
                    if (p > 0) GP[h * dd + 0 * d + j] = c0;
                    if (p > 1) GP[h * dd + 1 * d + j] = c1;
                }
            }
            if (q < 2) {
                tmem_ld2(tlane + GP2_COL, c0, c1);
                const int r = 128 + row, h = r / HP, j = r % HP;
                if (j < d) {
                    if (p > 0) GP[h * dd + 0 * d + j] = c0;
                    if (p > 1) GP[h * dd + 1 * d + j] = c1;
                }
            }
        }
        tc_fence_before();
        if (wt == 0) trace(17);
    }
    tmem_free_512(tmem, warp);
    if (threadIdx.x == 0) trace(18);
}

// ---------------------------------------------------------------------------------------------------------------
// kernel B3: dV~, dK~, per-head LayerNorm backward
// ---------------------------------------------------------------------------------------------------------------
struct KvBwdArgs {
    const float* gpart;         // (B, tiles, 4, d, d)
    const unsigned char* keep_mask;
    float mask_p, scale;
    unsigned long long mask_seed;
    const unsigned long long* seed_off;
    const float* vec;           // gamma / beta tables
    const float* pos;
    const float* qkv;           // (B n, 384): x^_K at 128.., x^_V at 256..
    const float* rstd_k;
    const float* rstd_v;
    float* dqkv;                // columns 128..383 written here
    float* part;
    int B, n, p, tiles, has_norm;
};
constexpr int KV_SMEM = 8 * TILE_BYTES + 2 * TILE_BYTES + 19 * 1024 + NWORK * STAGE_BYTES + 512 * 4 + 4 * 512 * 4 + 512 + 1024;
static_assert(KV_SMEM <= 232448 && AB_SMEM <= 232448, "shared memory budget");

__global__ void __launch_bounds__(THREADS, 1) enc_kv_bwd_kernel(const __grid_constant__ CUtensorMap mapQKV, KvBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_base(smem_raw);
    uint8_t* RK = smem;                                        // x^_K tile -> K~ operand
    uint8_t* RV = RK + 4 * TILE_BYTES;
    uint8_t* BOP = RV + 4 * TILE_BYTES;                        // dA operands: [v | k][head pair][hi 4 KB | lo 4 KB]
    float* dAs = reinterpret_cast<float*>(BOP + 2 * TILE_BYTES);       // [4][d][d]
    float* staging = dAs + 19 * 256;
    float* gb = staging + NWORK * 32 * STAGE_PITCH;            // gamma_K | beta_K | gamma_V | beta_V (4 x 128)
    float* csum = gb + 4 * 128;                                // [4 quarters][K|V][gamma|beta][128]
    Bars* bar = reinterpret_cast<Bars*>(csum + 4 * 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x % a.tiles;
    const int t0 = tile * TM;
    const int nvalid = min(TM, a.n - t0);
    const int p = a.p, d = DK + p, dd = d * d;

    if (threadIdx.x == 0) {
        mbar_init(&bar->xfull, 1);
        mbar_init(&bar->done, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&bar->dfull[i], 1); mbar_init(&bar->hand[i], NWORK); }
        partials_init(bar);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQKV) : "memory");
    }
    if (threadIdx.x == 0) { trace_entry(); trace(0); }
    const uint32_t tmem = tmem_alloc_512(&bar->tmem_slot, warp);
    grid_dep_sync();
    constexpr int DV_COL = 0, DK_COL = 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&bar->xfull, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(RK + j * TILE_BYTES, &mapQKV, &bar->xfull, 128 + 32 * j, t0, b);
            // tile partials of G stream through the V tile's (not yet needed) shared memory, then V lands there
            partials_produce(bar, RV, a.gpart + (long long)b * a.tiles * NH * dd, a.tiles, NH * dd);
            mbar_expect_tx(&bar->done, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(RV + j * TILE_BYTES, &mapQKV, &bar->done, 256 + 32 * j, t0, b);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t rk = smem_u32(RK), rv = smem_u32(RV), bo = smem_u32(BOP);
            const uint32_t id32 = idesc_bf16(32);
            mbar_wait(&bar->hand[0], 0);
            tc_fence_after();
            trace(4);
#pragma unroll
            for (int h = 0; h < 4; ++h)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t ko = (h & 1) * 64 + ks * 32;
                    const uint32_t kh = rk + (2 * (h >> 1)) * TILE_BYTES + ko, vh = rv + (2 * (h >> 1)) * TILE_BYTES + ko;
                    const uint32_t bv = bo + (h >> 1) * 8192 + ko, bk = bo + TILE_BYTES + (h >> 1) * 8192 + ko;
                    mma3(tmem + DV_COL + 32 * h, kh, kh + TILE_BYTES, bv, bv + 4096, id32, ks == 0);   // dV~ = K~ dA
                    mma3(tmem + DK_COL + 32 * h, vh, vh + TILE_BYTES, bk, bk + 4096, id32, ks == 0);   // dK~ = V~ dA^T
                }
            tc_commit(&bar->dfull[0]);
            trace(8);
        }
    } else {
        const int w = warp - 2, q = warp & 3, hf = w >> 2;
        const int wt = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const bool valid = row < nvalid;
        const long long grow = (long long)b * a.n + t0 + row;
        const long long grow0 = (long long)b * a.n + t0 + q * 32;
        const int nrows = max(0, min(32, nvalid - q * 32));
        float* stage = staging + w * 32 * STAGE_PITCH;
        float pv[2] = {0.f, 0.f};
        if (valid) {
            if (p > 0) pv[0] = a.pos[grow * p];
            if (p > 1) pv[1] = a.pos[grow * p + 1];
        }
        for (int i = wt; i < 512; i += NWORK * 32) gb[i] = a.has_norm ? a.vec[VEC_GK + i] : (((i >> 7) & 1) ? 0.f : 1.f);
        {
            unsigned long long mseed = a.mask_seed;
            if (a.mask_p > 0.f && a.seed_off) mseed += *a.seed_off;
            constexpr int EPT = (NH * 34 * 34 + NWORK * 32 - 1) / (NWORK * 32);
            float acc[EPT];
            const int ne = NH * dd;
            partials_consume<EPT>(bar, RV, a.tiles, ne, wt, lane, acc);
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int e = wt + i * NWORK * 32;
                if (e < ne) {
                    float sv = acc[i] * a.scale;
                    const long long E = (long long)b * ne + e;
                    if (a.keep_mask) sv *= 2.f * (float)a.keep_mask[E];
                    else if (a.mask_p > 0.f) sv *= dropout_scale(a.mask_p, mseed, (unsigned long long)E);
                    dAs[e] = sv;
                }
            }
        }
        worker_bar();
        if (wt == 0) trace(10);
        // B operands (K-major, 32 rows x 32 k per head; heads 2c, 2c+1 share image c at K offsets 0 / 32):
        //   dV~_h[t, j] = sum_i K~_h[t, i] dA_h[p+i][p+j]   ->  Bv[n = j][k = i]
        //   dK~_h[t, i] = sum_j V~_h[t, j] dA_h[p+i][p+j]   ->  Bk[n = i][k = j]
        for (int u = wt; u < 2 * 4 * 32 * 4; u += NWORK * 32) {
            const int which = u >> 9, h = (u >> 7) & 3, nrow = (u >> 2) & 31, uu = u & 3;
            float x8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                x8[k] = which == 0 ? dAs[h * dd + (p + 8 * uu + k) * d + p + nrow] : dAs[h * dd + (p + nrow) * d + p + 8 * uu + k];
            uint8_t* hi = BOP + which * TILE_BYTES + (h >> 1) * 8192;
            store_unit(hi, hi + 4096, nrow, (h & 1) * 4 + uu, x8);
        }
        mbar_wait(&bar->xfull, 0);
        split_tile_affine(RK, wt & 127, wt >> 7, a.has_norm ? gb : nullptr, gb + 128);
        mbar_wait(&bar->done, 0);
        split_tile_affine(RV, wt & 127, wt >> 7, a.has_norm ? gb + 256 : nullptr, gb + 384);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->hand[0]);
        if (wt == 0) trace(11);

        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        float v[32];
        mbar_wait(&bar->dfull[0], 0);
        tc_fence_after();
        if (wt == 0) trace(14);
        for (int blk = 0; blk < 2; ++blk) {                    // 0: V, 1: K
            for (int hh = 0; hh < 2; ++hh) {
                const int h = 2 * hf + hh;
                tmem_ld32(tlane + (blk == 0 ? DV_COL : DK_COL) + 32 * h, v);
                const float* Ah = dAs + h * dd;
                if (blk == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += pv[0] * Ah[p + j] + pv[1] * Ah[d + p + j];
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += pv[0] * Ah[(p + j) * d] + pv[1] * Ah[(p + j) * d + 1];
                }
                const int col = (blk == 0 ? 256 : 128) + 32 * h;
                float s1 = 0.f, s2 = 0.f;
                if (a.has_norm) {
                    float xh[32];
                    warp_load_block(stage, xh, lane, a.qkv + grow0 * 384 + col, 384, nrows);
                    const float rs = valid ? (blk == 0 ? a.rstd_v : a.rstd_k)[grow * NH + h] : 0.f;
                    const float* gam = gb + (blk == 0 ? 256 : 0) + 32 * h;
                    // column sums over this warp's 32 tokens through the staging block: d beta = sum dY, d gamma = sum dY x^
                    float t[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] = valid ? v[j] : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(&stage[lane * STAGE_PITCH + j]) = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
                    __syncwarp();
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) s2 += stage[r * STAGE_PITCH + lane];
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(&stage[lane * STAGE_PITCH + j]) =
                            make_float4(t[j] * xh[j], t[j + 1] * xh[j + 1], t[j + 2] * xh[j + 2], t[j + 3] * xh[j + 3]);
                    __syncwarp();
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) s1 += stage[r * STAGE_PITCH + lane];
                    __syncwarp();
                    float m1 = 0.f, m2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] *= gam[j];
                        m1 += v[j];
                        m2 = fmaf(v[j], xh[j], m2);
                    }
                    m1 *= (1.f / 32.f);
                    m2 *= (1.f / 32.f);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = rs * (v[j] - m1 - xh[j] * m2);
                }
                csum[q * 512 + (1 - blk) * 256 + 32 * h + lane] = s1;              // [K|V][gamma]
                csum[q * 512 + (1 - blk) * 256 + 128 + 32 * h + lane] = s2;        // [K|V][beta]
                warp_store_block(stage, v, lane, a.dqkv + grow0 * 384 + col, 384, nrows);
            }
        }
        tc_fence_before();
        worker_bar();
        float* P = a.part + (long long)(b * a.tiles + tile) * VEC_FLOATS;
        for (int i = wt; i < 512; i += NWORK * 32)             // VEC order: gamma_K | beta_K | gamma_V | beta_V
            P[VEC_GK + i] = (csum[i] + csum[512 + i]) + (csum[1024 + i] + csum[1536 + i]);
        if (wt == 0) trace(17);
    }
    tmem_free_512(tmem, warp);
    if (threadIdx.x == 0) trace(18);
}

// ---------------------------------------------------------------------------------------------------------------
// kernel B4: dx = dqkv W_qkv + dx1, column sums of dqkv
// ---------------------------------------------------------------------------------------------------------------
struct DxArgs {
    const uint8_t* qkvt;        // 12 tiles of W_qkv^T: (K chunk 0..5 of the 384-wide index, hi/lo)
    const float* dx1;           // residual branch gradient (B n, 128) or null
    float* dx;
    float* part;
    int B, n, tiles;
};
constexpr int DX_RING = 2;
constexpr int DX_SMEM = 8 * TILE_BYTES + DX_RING * TILE_BYTES + TILE_BYTES + NWORK * STAGE_BYTES + 512 + 1024;

struct DxBars {
    uint64_t afull[3], aconv[3], afree, full[DX_RING], empty[DX_RING], ready, dfull;
    uint32_t tmem_slot;
};

__global__ void __launch_bounds__(THREADS, 1) enc_dx_kernel(const __grid_constant__ CUtensorMap mapG, DxArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_base(smem_raw);
    uint8_t* AB = smem;                                        // two A buffers of 4 tiles
    uint8_t* ring = AB + 8 * TILE_BYTES;
    uint8_t* ones = ring + DX_RING * TILE_BYTES;
    float* staging = reinterpret_cast<float*>(ones + TILE_BYTES);
    DxBars* bar = reinterpret_cast<DxBars*>(staging + NWORK * 32 * STAGE_PITCH);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x % a.tiles;
    const int t0 = tile * TM;
    const int nvalid = min(TM, a.n - t0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&bar->afull[i], 1); mbar_init(&bar->aconv[i], NWORK); }
        mbar_init(&bar->afree, 1);
        mbar_init(&bar->ready, NWORK);
        mbar_init(&bar->dfull, 1);
        for (int s = 0; s < DX_RING; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapG) : "memory");
    }
    if (threadIdx.x == 0) { trace_entry(); trace(0); }
    const uint32_t tmem = tmem_alloc_512(&bar->tmem_slot, warp);
    grid_dep_sync();
    constexpr int DB_COL = 128;

    if (warp == 0) {
        if (lane == 0) {
            auto load_a = [&](int blk) {
                uint8_t* dst = AB + (blk & 1) * 4 * TILE_BYTES;
                mbar_expect_tx(&bar->afull[blk], 4 * TILE_BYTES);
                for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * TILE_BYTES, &mapG, &bar->afull[blk], blk * 128 + 32 * j, t0, b);
            };
            load_a(0);
            load_a(1);
            for (int i = 0; i < 12; ++i) {
                if (i == 4) {                                  // third block reuses buffer 0 once block 0's MMAs are done
                    mbar_wait(&bar->afree, 0);
                    load_a(2);
                }
                const int s = i % DX_RING;
                mbar_wait(&bar->empty[s], ((i / DX_RING) & 1) ^ 1);
                mbar_expect_tx(&bar->full[s], TILE_BYTES);
                bulk_load(ring + s * TILE_BYTES, a.qkvt + (size_t)i * TILE_BYTES, TILE_BYTES, &bar->full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ab = smem_u32(AB), rb = smem_u32(ring), on = smem_u32(ones);
            const uint32_t id128 = idesc_bf16(128);
            mbar_wait(&bar->ready, 0);
            for (int blk = 0; blk < 3; ++blk) {
                const uint32_t abuf = ab + (blk & 1) * 4 * TILE_BYTES;
                mbar_wait(&bar->aconv[blk], 0);
                tc_fence_after();
                mma_colsum(tmem + DB_COL + 16 * blk, abuf, abuf + TILE_BYTES, 2 * TILE_BYTES, on);
                for (int i = 4 * blk; i < 4 * blk + 4; ++i) {
                    const int s = i % DX_RING;
                    mbar_wait(&bar->full[s], (i / DX_RING) & 1);
                    tc_fence_after();
                    const int kc = (i >> 1) & 1;
                    mma_weight_tile(tmem, abuf + (2 * kc) * TILE_BYTES, abuf + (2 * kc + 1) * TILE_BYTES, rb + s * TILE_BYTES,
                                    (i & 1) == 0, id128, blk == 0 && kc == 0);
                    tc_commit(&bar->empty[s]);
                }
                if (blk == 0) tc_commit(&bar->afree);
            }
            tc_commit(&bar->dfull);
            trace(8);
        }
    } else {
        const int w = warp - 2, q = warp & 3, hf = w >> 2;
        const int wt = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const long long grow0 = (long long)b * a.n + t0 + q * 32;
        const int nrows = max(0, min(32, nvalid - q * 32));
        float* stage = staging + w * 32 * STAGE_PITCH;
        write_ones(ones, wt);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->ready);
        for (int blk = 0; blk < 3; ++blk) {
            mbar_wait(&bar->afull[blk], 0);
            split_tile_inplace(AB + (blk & 1) * 4 * TILE_BYTES, wt & 127, wt >> 7);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar->aconv[blk]);
        }
        if (wt == 0) trace(11);
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        float v[32];
        mbar_wait(&bar->dfull, 0);
        tc_fence_after();
        if (wt == 0) trace(14);
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = hf * 64 + cc * 32;
            tmem_ld32(tlane + c0, v);
            if (a.dx1) {
                float r[32];
                warp_load_block(stage, r, lane, a.dx1 + grow0 * DM + c0, DM, nrows);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += r[j];
            }
            warp_store_block(stage, v, lane, a.dx + grow0 * DM + c0, DM, nrows);
        }
        float* P = a.part + (long long)(b * a.tiles + tile) * VEC_FLOATS;
        float s0, s1;
        if (hf == 0) {
            tmem_ld2(tlane + DB_COL, s0, s1);
            P[VEC_BQKV + row] = s0;
            tmem_ld2(tlane + DB_COL + 32, s0, s1);
            P[VEC_BQKV + 256 + row] = s0;
        } else {
            tmem_ld2(tlane + DB_COL + 16, s0, s1);
            P[VEC_BQKV + 128 + row] = s0;
        }
        tc_fence_before();
        if (wt == 0) trace(17);
    }
    tmem_free_512(tmem, warp);
    if (threadIdx.x == 0) trace(18);
}

// out[i] = sum over tiles of part[t][i], fixed order
__global__ void enc_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= VEC_FLOATS) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int t = 0;
    for (; t + 3 < nparts; t += 4) {
        s0 += part[(long long)t * VEC_FLOATS + i];
        s1 += part[(long long)(t + 1) * VEC_FLOATS + i];
        s2 += part[(long long)(t + 2) * VEC_FLOATS + i];
        s3 += part[(long long)(t + 3) * VEC_FLOATS + i];
    }
    for (; t < nparts; ++t) s0 += part[(long long)t * VEC_FLOATS + i];
    out[i] = (s0 + s1) + (s2 + s3);
}

}  // namespace enc
}  // namespace gb200

using namespace gb200;
using namespace gb200::enc;

extern "C" int gb200_encoder_bwd_set_trace(unsigned long long* device_buffer) {
    return cudaMemcpyToSymbol(g_enc_trace, &device_buffer, sizeof(device_buffer)) == cudaSuccess ? GB200_OK : GB200_ERR_CUDA;
}

extern "C" size_t gb200_encoder_bwd_workspace_bytes(int B, int n, int n_head, int d_k, int pos_dim) {
    const int tiles = (n + TM - 1) / TM, d = d_k + pos_dim;
    return ((size_t)B * tiles * n_head * d * d + (size_t)B * tiles * VEC_FLOATS) * sizeof(float);
}

extern "C" int gb200_encoder_layer_bwd(int device, const void* packed, int d_model, int n_head, int pos_dim, int d_ff,
                                       const float* dy, const float* pos, int B, int n, int has_norm, float attn_scale,
                                       const unsigned char* keep_mask, float mask_p, unsigned long long mask_seed,
                                       float p_attn_out, unsigned long long seed_attn_out, float res_sign, float p_ffn,
                                       float p_out, unsigned long long seed_out, const float* qkv, const float* rstd_k,
                                       const float* rstd_v, const float* attn, const float* hidden, float* g2, float* g1,
                                       float* dx1, float* gfc, float* dqkv, float* dx, float* dvec, float* workspace,
                                       size_t workspace_bytes, int stages, void* stream) {
    use_device(device);
    GB_REQUIRE(gb200_encoder_supported(d_model, n_head, pos_dim, d_ff), "gb200_encoder_layer_bwd: unsupported layer shape");
    GB_REQUIRE(packed && dy && pos && qkv && attn && hidden && g1 && dx1 && dqkv && dx && dvec && workspace,
               "gb200_encoder_layer_bwd: null buffer");
    GB_REQUIRE(!has_norm || (rstd_k && rstd_v), "gb200_encoder_layer_bwd: null rstd");
    GB_REQUIRE(p_out <= 0.f || g2, "gb200_encoder_layer_bwd: g2 buffer needed when the output dropout is on");
    GB_REQUIRE((p_attn_out <= 0.f && res_sign == 1.f) || gfc, "gb200_encoder_layer_bwd: gfc buffer needed");
    GB_REQUIRE(workspace_bytes >= gb200_encoder_bwd_workspace_bytes(B, n, n_head, DK, pos_dim),
               "gb200_encoder_layer_bwd: workspace too small");
    static bool configured = false;
    if (!configured) {
        set_smem(enc_ffn_bwd_kernel, FB_SMEM);
        set_smem(enc_attn_bwd_kernel, AB_SMEM);
        set_smem(enc_kv_bwd_kernel, KV_SMEM);
        set_smem(enc_dx_kernel, DX_SMEM);
        configured = true;
    }
    const int tiles = (n + TM - 1) / TM, d = DK + pos_dim;
    const uint8_t* wt = reinterpret_cast<const uint8_t*>(packed);
    const float* vec = reinterpret_cast<const float*>(wt + (size_t)GB200_ENC_TILES * TILE_BYTES);
    float* gpart = workspace;
    float* part = workspace + (size_t)B * tiles * n_head * d * d;
    cudaStream_t st = as_stream(stream);
    const bool need_gfc = p_attn_out > 0.f || res_sign != 1.f;
    int launched = 0;
    if (stages & 1) {
        CUtensorMap m;
        GB_REQUIRE(make_tile_map(&m, dy, DM, n, B), "gb200_encoder_layer_bwd: tensor map (dy) failed");
        FfnBwdArgs a;
        a.w2t = wt + (size_t)TS_W2T * TILE_BYTES; a.w1t = wt + (size_t)TS_W1T * TILE_BYTES; a.dy = dy; a.hid = hidden;
        a.g2 = p_out > 0.f ? g2 : nullptr; a.g1 = g1; a.dx1 = dx1; a.part = part; a.p2 = p_out; a.pf = p_ffn; a.seed2 = seed_out;
        a.seed_off = rng_offset_ptr(); a.B = B; a.n = n; a.tiles = tiles;
        launch_enc(enc_ffn_bwd_kernel, B * tiles, FB_SMEM, st, m, a);
        ++launched;
    }
    if (stages & 2) {
        CUtensorMap mdx, mq;
        GB_REQUIRE(make_tile_map(&mdx, dx1, DM, n, B) && make_tile_map(&mq, qkv, 3 * DM, n, B),
                   "gb200_encoder_layer_bwd: tensor map (dx1 / qkv) failed");
        AttnBwdArgs a;
        a.fct = wt + (size_t)TS_FCT * TILE_BYTES; a.dx1 = dx1; a.gfc = need_gfc ? gfc : nullptr; a.p1 = p_attn_out; a.sign = res_sign;
        a.seed1 = seed_attn_out; a.seed_off = rng_offset_ptr(); a.attn = attn; a.pos = pos; a.dqkv = dqkv; a.gpart = gpart;
        a.part = part; a.B = B; a.n = n; a.p = pos_dim; a.tiles = tiles;
        launch_enc(enc_attn_bwd_kernel, B * tiles, AB_SMEM, st, mdx, mq, a);
        ++launched;
    }
    if (stages & 4) {
        CUtensorMap mq;
        GB_REQUIRE(make_tile_map(&mq, qkv, 3 * DM, n, B), "gb200_encoder_layer_bwd: tensor map (qkv) failed");
        KvBwdArgs a;
        a.gpart = gpart; a.keep_mask = keep_mask; a.mask_p = keep_mask ? 0.f : mask_p; a.scale = attn_scale; a.mask_seed = mask_seed;
        a.seed_off = rng_offset_ptr(); a.vec = vec; a.pos = pos; a.qkv = qkv; a.rstd_k = rstd_k; a.rstd_v = rstd_v; a.dqkv = dqkv;
        a.part = part; a.B = B; a.n = n; a.p = pos_dim; a.tiles = tiles; a.has_norm = has_norm;
        launch_enc(enc_kv_bwd_kernel, B * tiles, KV_SMEM, st, mq, a);
        ++launched;
    }
    if (stages & 8) {
        CUtensorMap mg;
        GB_REQUIRE(make_tile_map(&mg, dqkv, 3 * DM, n, B), "gb200_encoder_layer_bwd: tensor map (dqkv) failed");
        DxArgs a;
        a.qkvt = wt + (size_t)TS_QKVT * TILE_BYTES; a.dx1 = dx1; a.dx = dx; a.part = part; a.B = B; a.n = n; a.tiles = tiles;
        launch_enc(enc_dx_kernel, B * tiles, DX_SMEM, st, mg, a);
        ++launched;
    }
    if (stages & 16) {
        enc_reduce_kernel<<<(VEC_FLOATS + 127) / 128, 128, 0, st>>>(part, B * tiles, dvec);
        ++launched;
    }
    return check_launch("gb200_encoder_layer_bwd", launched);
}
