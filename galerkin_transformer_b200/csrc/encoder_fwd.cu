// Fused forward of the Galerkin encoder layer (libs/model.py:104-140 with libs/layers.py:829-899, 979-987) in three
// tcgen05 kernels; CTA = one 128-token tile of one sample (tiles never straddle samples):
//
//   enc_qkv_kernel   x -> [Q | LN(K) | LN(V)] (+rstd)  and the tile's partial  K~^T V~   (the reduction over tokens
//                    is finished, scaled and masked by the next kernel)                    layers.py:837-851, 869-874, 723
//   enc_attn_kernel  A = mask * sum(partials) / n ;  heads = Q~ A ;  x1 = x +/- drop(heads W_fc^T + b)
//                                                                                          layers.py:728-733, 892-897, model.py:124-127
//   enc_ffn_kernel   x2 = x1 + drop( drop(relu(x1 W1^T + b1)) W2^T + b2 )                  layers.py:979-987, model.py:131-132
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warps 2-9 = workers
// (fp32 -> bf16 split of landed tiles, TMEM epilogues, operand hand-over to the next MMA through shared memory).
// All GEMMs are bf16x3 (encoder_common.cuh).  Specialised to d_model = 128, 4 heads of 32, d_ff = 256, pos_dim <= 2.
#include <cuda.h>

#include "encoder_common.cuh"

namespace gb200 {
namespace enc {

// ---------------------------------------------------------------------------------------------------------------
// kernel 1: Q|K|V projection + per-head LayerNorm + partial K~^T V~
// ---------------------------------------------------------------------------------------------------------------
struct QkvArgs {
    const uint8_t* wtiles;      // 12 tiles: (block Q/K/V, K chunk 0/1, hi/lo)
    const float* vec;           // bias / LayerNorm tables (VEC_* offsets)
    const float* pos;           // (B, n, p)
    float* qkv;                 // (B n, 384): Q | x^_K | x^_V
    float* rstd_k;              // (B n, 4)
    float* rstd_v;
    float* part;                // (B, tiles, 4, d, d) partial K~^T V~
    int B, n, p, tiles, has_norm;
    float eps;
};

constexpr int QKV_RING = 6;
constexpr int QKV_SMEM = 4 * TILE_BYTES + QKV_RING * TILE_BYTES + NWORK * STAGE_BYTES + VEC_FLOATS * 4 + 512 + 1024;

__global__ void __launch_bounds__(THREADS, 1) enc_qkv_kernel(const __grid_constant__ CUtensorMap mapX, QkvArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_base(smem_raw);
    uint8_t* X = smem;                                        // x tile, later the K~^T operand (A2)
    uint8_t* ring = smem + 4 * TILE_BYTES;                    // weight tiles, later the [V~ | pos]^T operand (B2)
    float* staging = reinterpret_cast<float*>(ring + QKV_RING * TILE_BYTES);
    float* vec = staging + NWORK * 32 * STAGE_PITCH;
    float* ppw = vec + VEC_FLOATS;                            // [4 warps][4] partial pos^T pos
    Bars* bar = reinterpret_cast<Bars*>(ppw + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x % a.tiles;
    const int t0 = tile * TM;
    const int nvalid = min(TM, a.n - t0);
    const int p = a.p, d = DK + p;

    if (threadIdx.x == 0) {
        mbar_init(&bar->xfull, 1);
        mbar_init(&bar->xconv, NWORK);
        for (int s = 0; s < QKV_RING; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar->dfull[i], 1); mbar_init(&bar->hand[i], NWORK); }
        mbar_init(&bar->done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    }
    if (threadIdx.x == 0) { trace_entry(); trace(0); }
    const uint32_t tmem = tmem_alloc_512(&bar->tmem_slot, warp);
    grid_dep_sync();
    if (threadIdx.x == 0) trace(1);

    if (warp == 0) {
        if (lane == 0) {   // ---------------- producer ----------------
            mbar_expect_tx(&bar->xfull, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(X + j * TILE_BYTES, &mapX, &bar->xfull, 32 * j, t0, b);
            for (int i = 0; i < 12; ++i) {
                const int s = i % QKV_RING;
                mbar_wait(&bar->empty[s], ((i / QKV_RING) & 1) ^ 1);
                mbar_expect_tx(&bar->full[s], TILE_BYTES);
                bulk_load(ring + s * TILE_BYTES, a.wtiles + (size_t)i * TILE_BYTES, TILE_BYTES, &bar->full[s]);
                if (i == 5) trace(2);
            }
            trace(3);
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---------------- MMA issuer ----------------
            const uint32_t xa = smem_u32(X), rb = smem_u32(ring);
            const uint32_t id128 = idesc_bf16(128);
            mbar_wait(&bar->xconv, 0);
            tc_fence_after();
            trace(4);
            for (int i = 0; i < 12; ++i) {
                const int s = i % QKV_RING;
                mbar_wait(&bar->full[s], (i / QKV_RING) & 1);
                tc_fence_after();
                if (i == 0) trace(5);
                const int nb = i >> 2, kc = (i >> 1) & 1;
                mma_weight_tile(tmem + nb * 128, xa + (2 * kc) * TILE_BYTES, xa + (2 * kc + 1) * TILE_BYTES,
                                rb + s * TILE_BYTES, (i & 1) == 0, id128, kc == 0);
                tc_commit(&bar->empty[s]);
                if ((i & 3) == 3) tc_commit(&bar->dfull[nb]);
            }
            // second contraction, over the tile's tokens:  D2 = K~^T [V~ | pos]   (M = 128 K features, N = 144),
            //                                              D3 = V~^T pos          (M = 128 V features, N = 16)
            trace(6);
            mbar_wait(&bar->hand[0], 0);
            tc_fence_after();
            trace(7);
            // operands are MN-major chunk images [128 token rows][64 features]: K~ = X blocks {0: hi f<64, 1: lo f<64,
            // 2: hi f>=64, 3: lo f>=64}; [V~ | pos] = ring blocks {hi, lo} x {f<64, f>=64, pos}
            const uint32_t id144 = idesc_bf16_mn(144), id16 = idesc_bf16_mn(16);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t ah = xa + ks * 2048, al = ah + TILE_BYTES;
                const uint32_t bh = rb + ks * 2048, bl = bh + TILE_BYTES;
                mma3_mn(tmem, ah, al, 2 * TILE_BYTES, bh, bl, 2 * TILE_BYTES, id144, ks == 0);
                mma3_mn(tmem + 144, bh, bl, 2 * TILE_BYTES, bh + 4 * TILE_BYTES, bl + 4 * TILE_BYTES, 2 * TILE_BYTES, id16,
                        ks == 0);
            }
            tc_commit(&bar->dfull[3]);
            trace(8);
        }
    } else {               // ---------------- workers ----------------
        const int w = warp - 2, q = warp & 3, hf = w >> 2;
        const int wt = threadIdx.x - 64;                       // 0..255
        const int row = q * 32 + lane;                         // tile row == TMEM lane
        const bool valid = row < nvalid;
        const long long grow = (long long)b * a.n + t0 + row;
        const int nrows = max(0, min(32, nvalid - q * 32));
        float* stage = staging + w * 32 * STAGE_PITCH;
        for (int i = wt; i < VEC_FLOATS; i += NWORK * 32) vec[i] = a.vec[i];
        float pv[2] = {0.f, 0.f};                              // this row's position (needed late: fetched early)
        if (valid && hf == 0) {
            if (p > 0) pv[0] = a.pos[grow * p];
            if (p > 1) pv[1] = a.pos[grow * p + 1];
        }
        mbar_wait(&bar->xfull, 0);
        if (wt == 0) trace(9);
        split_tile_inplace(X, wt & 127, wt >> 7);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->xconv);
        worker_bar();                                          // vec[] visible to every worker
        if (wt == 0) trace(10);

        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        float v[32];
        // Q chunks of heads 2hf, 2hf+1: bias, store
        mbar_wait(&bar->dfull[0], 0);
        tc_fence_after();
        if (wt == 0) trace(11);
        for (int hh = 0; hh < 2; ++hh) {
            const int h = 2 * hf + hh;
            tmem_ld32(tlane + h * 32, v);
            const float4* bb = reinterpret_cast<const float4*>(vec + VEC_BQKV + h * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = bb[j];
                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
            warp_store_block(stage, v, lane, a.qkv + ((long long)b * a.n + t0 + q * 32) * 384 + h * 32, 384, nrows);
        }
        // every projection MMA has completed: the x tile and the weight ring may be overwritten
        if (wt == 0) trace(12);
        mbar_wait(&bar->dfull[2], 0);
        tc_fence_after();
        if (wt == 0) trace(13);
        uint8_t* B2 = ring;
        for (int blk = 1; blk <= 2; ++blk) {                   // 1: K -> X blocks, 2: V -> ring blocks
            for (int hh = 0; hh < 2; ++hh) {
                const int h = 2 * hf + hh;
                tmem_ld32(tlane + blk * 128 + h * 32, v);
                {
                    const float4* bb = reinterpret_cast<const float4*>(vec + VEC_BQKV + blk * 128 + h * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = bb[j];
                        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                    }
                }
                if (a.has_norm) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) { s0 += v[j]; s1 += v[j + 1]; s2 += v[j + 2]; s3 += v[j + 3]; }
                    const float mean = ((s0 + s1) + (s2 + s3)) * (1.f / 32.f);
                    s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        v[j] -= mean; v[j + 1] -= mean; v[j + 2] -= mean; v[j + 3] -= mean;
                        s0 = fmaf(v[j], v[j], s0); s1 = fmaf(v[j + 1], v[j + 1], s1);
                        s2 = fmaf(v[j + 2], v[j + 2], s2); s3 = fmaf(v[j + 3], v[j + 3], s3);
                    }
                    const float rs = rsqrtf(((s0 + s1) + (s2 + s3)) * (1.f / 32.f) + a.eps);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= rs;
                    if (valid) (blk == 1 ? a.rstd_k : a.rstd_v)[grow * NH + h] = rs;
                }
                warp_store_block(stage, v, lane, a.qkv + ((long long)b * a.n + t0 + q * 32) * 384 + blk * 128 + h * 32,
                                 384, nrows);
                // operand of the token contraction: affine applied, rows past the sample's end are zero.
                // MN-major image: row = token, head h = 4 units of MN block h / 2
                if (a.has_norm) {
                    const float4* gg = reinterpret_cast<const float4*>(vec + (blk == 1 ? VEC_GK : VEC_GV) + h * 32);
                    const float4* be = reinterpret_cast<const float4*>(vec + (blk == 1 ? VEC_BK : VEC_BV) + h * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 g4 = gg[j], b4 = be[j];
                        v[4 * j] = fmaf(v[4 * j], g4.x, b4.x); v[4 * j + 1] = fmaf(v[4 * j + 1], g4.y, b4.y);
                        v[4 * j + 2] = fmaf(v[4 * j + 2], g4.z, b4.z); v[4 * j + 3] = fmaf(v[4 * j + 3], g4.w, b4.w);
                    }
                }
                if (!valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                uint8_t* hi = (blk == 1 ? X : B2) + (h >> 1) * 2 * TILE_BYTES;
#pragma unroll
                for (int u = 0; u < 4; ++u) store_unit(hi, hi + TILE_BYTES, row, (h & 1) * 4 + u, &v[8 * u]);
                if (wt == 0) trace(21 + (blk - 1) * 2 + hh);
            }
        }
        if (wt == 0) trace(20);
        if (hf == 0) {     // position block of B2 (MN block 2: 16 columns, zero beyond p) and this warp's share of pos^T pos
            float x8[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x8[c] = 0.f;
            x8[0] = pv[0]; x8[1] = pv[1];
            uint8_t* hi = B2 + 4 * TILE_BYTES;
            store_unit(hi, hi + TILE_BYTES, row, 0, &x8[0]);
            store_unit(hi, hi + TILE_BYTES, row, 1, &x8[8]);
            const float p00 = warp_sum(pv[0] * pv[0]), p01 = warp_sum(pv[0] * pv[1]), p11 = warp_sum(pv[1] * pv[1]);
            if (lane == 0) { ppw[q * 4 + 0] = p00; ppw[q * 4 + 1] = p01; ppw[q * 4 + 2] = p01; ppw[q * 4 + 3] = p11; }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->hand[0]);
        if (wt == 0) trace(14);

        // partial attention matrix of this tile: A_h[i][j], i/j = 0..p-1 position, p.. = feature
        mbar_wait(&bar->dfull[3], 0);
        tc_fence_after();
        worker_bar();                                          // ppw[] complete
        if (wt == 0) trace(15);
        float* P = a.part + ((long long)(b * a.tiles + tile) * NH + q) * d * d;       // head q
        if (hf == 0) {     // rows = K features of head q: columns [V features of head q | pos]
            tmem_ld32(tlane + q * 32, v);
            float c0, c1;
            tmem_ld2(tlane + 128, c0, c1);
            float* prow = P + (p + lane) * d;
            if (p > 0) prow[0] = c0;
            if (p > 1) prow[1] = c1;
#pragma unroll
            for (int j = 0; j < 32; ++j) prow[p + j] = v[j];
        } else {           // rows = V features of head q: D3[j][c] = sum_t V~[t,j] pos[t,c]  ->  A_h[c][p + j]
            float c0, c1;
            tmem_ld2(tlane + 144, c0, c1);
            if (p > 0) P[0 * d + p + lane] = c0;
            if (p > 1) P[1 * d + p + lane] = c1;
            if (lane < p * p) {
                const int ci = lane / p, cj = lane % p;
                const int e = ci * 2 + cj;
                P[ci * d + cj] = ppw[e] + ppw[4 + e] + ppw[8 + e] + ppw[12 + e];
            }
        }
        tc_fence_before();
        if (wt == 0) trace(16);
    }
    tmem_free_512(tmem, warp);
    if (threadIdx.x == 0) trace(17);
}

// ---------------------------------------------------------------------------------------------------------------
// kernel 2: attention matrix reduction, heads = Q~ A, x1 = x + sign * drop(heads W_fc^T + b_fc)
// ---------------------------------------------------------------------------------------------------------------
struct AttnArgs {
    const uint8_t* wtiles;      // 6 tiles of W_fc' (K chunk 0..2, hi/lo), head-padded K layout (h * 48 + j)
    const float* vec;
    const float* pos;
    const float* part;          // (B, tiles, 4, d, d)
    const unsigned char* keep_mask;   // (B, 4, d, d) or null
    float mask_p;
    unsigned long long mask_seed;
    const unsigned long long* seed_off;
    float scale;
    float* attn;                // (B, 4, d, d): masked, scaled attention matrix (module output, saved for backward)
    float* heads;               // (B n, 4 d)
    const float* x;             // residual (B n, 128)
    float* x1;
    float p1;
    unsigned long long seed1;
    float sign;
    int B, n, p, tiles;
};
constexpr int ATT_RING = 3;
constexpr int ATT_R_BYTES = 6 * TILE_BYTES;                       // Q tile (4) + B_a, later the heads operand (3 x hi/lo)
constexpr int ATT_AS_BYTES = 4 * 34 * 34 * 4 + 192;               // 18688: fp32 A for the rank-p update (multiple of 128)
constexpr int ATT_SMEM = ATT_R_BYTES + 19 * 1024 + ATT_RING * TILE_BYTES + NWORK * STAGE_BYTES + 512 + 512 + 1024;

__global__ void __launch_bounds__(THREADS, 1) enc_attn_kernel(const __grid_constant__ CUtensorMap mapQ, AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_base(smem_raw);
    uint8_t* R = smem;                                         // Q tile | B_a  ->  heads operand
    uint8_t* BA = R + 4 * TILE_BYTES;                          // 2 chunks x (hi, lo) x [48 rows x 128 B]
    float* As = reinterpret_cast<float*>(R + ATT_R_BYTES);     // [4][d][d]
    uint8_t* ring = R + ATT_R_BYTES + 19 * 1024;
    float* staging = reinterpret_cast<float*>(ring + ATT_RING * TILE_BYTES);
    float* bfc = staging + NWORK * 32 * STAGE_PITCH;           // [128]
    Bars* bar = reinterpret_cast<Bars*>(bfc + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x % a.tiles;
    const int t0 = tile * TM;
    const int nvalid = min(TM, a.n - t0);
    const int p = a.p, d = DK + p, dd = d * d;

    if (threadIdx.x == 0) {
        mbar_init(&bar->xfull, 1);
        mbar_init(&bar->xconv, NWORK);
        for (int s = 0; s < ATT_RING; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar->dfull[i], 1); mbar_init(&bar->hand[i], NWORK); }
        partials_init(bar);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQ) : "memory");
    }
    if (threadIdx.x == 0) { trace_entry(); trace(0); }
    const uint32_t tmem = tmem_alloc_512(&bar->tmem_slot, warp);
    grid_dep_sync();
    if (threadIdx.x == 0) trace(1);
    constexpr uint32_t BAC = HP * 128;                         // one [48 x 64] chunk image of B_a
    constexpr int DB_COL = 4 * HP;                             // fc accumulator columns 192..319

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&bar->xfull, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(R + j * TILE_BYTES, &mapQ, &bar->xfull, 32 * j, t0, b);
            // tile partials of the attention matrix through the (still idle) weight ring + staging area
            partials_produce(bar, ring, a.part + (long long)b * a.tiles * NH * dd, a.tiles, NH * dd);
            for (int i = 0; i < 6; ++i) {
                const int s = i % ATT_RING;
                mbar_wait(&bar->empty[s], ((i / ATT_RING) & 1) ^ 1);
                mbar_expect_tx(&bar->full[s], TILE_BYTES);
                bulk_load(ring + s * TILE_BYTES, a.wtiles + (size_t)i * TILE_BYTES, TILE_BYTES, &bar->full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ra = smem_u32(R), ba = smem_u32(BA), rb = smem_u32(ring);
            const uint32_t id48 = idesc_bf16(HP), id128 = idesc_bf16(128);
            mbar_wait(&bar->xconv, 0);
            mbar_wait(&bar->hand[0], 0);
            tc_fence_after();
            trace(4);
            // heads_h = Q_h A_h[p:, :]  (the rank-p position part is added by the epilogue): K = 32 = two k-steps
#pragma unroll
            for (int h = 0; h < 4; ++h)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t ko = (h & 1) * 64 + ks * 32;
                    const uint32_t ah = ra + (2 * (h >> 1)) * TILE_BYTES + ko, al = ah + TILE_BYTES;
                    const uint32_t bh = ba + (2 * (h >> 1)) * BAC + ko, bl = bh + BAC;
                    mma3(tmem + h * HP, ah, al, bh, bl, id48, ks == 0);
                }
            tc_commit(&bar->dfull[0]);
            mbar_wait(&bar->hand[1], 0);
            tc_fence_after();
            trace(7);
            for (int i = 0; i < 6; ++i) {
                const int s = i % ATT_RING;
                mbar_wait(&bar->full[s], (i / ATT_RING) & 1);
                tc_fence_after();
                const int kc = i >> 1;
                mma_weight_tile(tmem + DB_COL, ra + (2 * kc) * TILE_BYTES, ra + (2 * kc + 1) * TILE_BYTES,
                                rb + s * TILE_BYTES, (i & 1) == 0, id128, kc == 0);
                tc_commit(&bar->empty[s]);
            }
            tc_commit(&bar->dfull[1]);
            trace(8);
        }
    } else {
        const int w = warp - 2, q = warp & 3, hf = w >> 2;
        const int wt = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const bool valid = row < nvalid;
        if (wt == 0) trace(9);
        const long long grow = (long long)b * a.n + t0 + row;
        const long long grow0 = (long long)b * a.n + t0 + q * 32;
        const int nrows = max(0, min(32, nvalid - q * 32));
        float* stage = staging + w * 32 * STAGE_PITCH;
        if (wt < 128) bfc[wt] = a.vec[VEC_BFC + wt];
        // attention matrix: fixed-order sum of the tile partials, 1/n (or 1/(sqrt(d) n)), dropout mask
        {
            unsigned long long mseed = a.mask_seed;
            if (a.mask_p > 0.f && a.seed_off) mseed += *a.seed_off;
            // every thread owns <= 19 elements; the tile partials arrive slot by slot in tile order (deterministic sum)
            constexpr int EPT = (NH * 34 * 34 + NWORK * 32 - 1) / (NWORK * 32);
            float acc[EPT];
            const int ne = NH * dd;
            partials_consume<EPT>(bar, ring, a.tiles, ne, wt, lane, acc);
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int e = wt + i * NWORK * 32;
                if (e < ne) {
                    float s = acc[i] * a.scale;
                    const long long E = (long long)b * ne + e;
                    if (a.keep_mask) s *= 2.f * (float)a.keep_mask[E];
                    else if (a.mask_p > 0.f) s *= dropout_scale(a.mask_p, mseed, (unsigned long long)E);
                    As[e] = s;
                    if (tile == 0) a.attn[E] = s;
                }
            }
        }
        worker_bar();
        if (wt == 0) trace(10);
        // B operand of heads_h = Q_h A_h[p:, :]:  row = output column j (48, zero past d), K = feature i of head h;
        // heads (2c, 2c+1) share chunk c (K offsets 0 and 32)
        for (int u = wt; u < 4 * HP * 4; u += NWORK * 32) {
            const int h = u / (HP * 4), j = (u / 4) % HP, i0 = (u & 3) * 8;
            float x8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x8[i] = (j < d) ? As[h * dd + (p + i0 + i) * d + j] : 0.f;
            uint8_t* hi = BA + (2 * (h >> 1)) * BAC;
            store_unit(hi, hi + BAC, j, (h & 1) * 4 + (u & 3), x8);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->hand[0]);
        if (wt == 0) trace(11);
        mbar_wait(&bar->xfull, 0);
        if (wt == 0) trace(12);
        split_tile_inplace(R, wt & 127, wt >> 7);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->xconv);
        if (wt == 0) trace(13);

        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        float v[32];
        float pv[2] = {0.f, 0.f};
        if (valid)
            for (int c = 0; c < p; ++c) pv[c] = a.pos[grow * p + c];
        mbar_wait(&bar->dfull[0], 0);                          // heads accumulators ready; Q tile and B_a are dead
        tc_fence_after();
        if (wt == 0) trace(14);
        for (int hh = 0; hh < 2; ++hh) {
            const int h = 2 * hf + hh;
            float e0, e1;
            tmem_ld32(tlane + h * HP, v);
            tmem_ld2(tlane + h * HP + 32, e0, e1);
            const float* A0 = As + h * dd;                     // position rows of A_h
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += pv[0] * A0[j] + pv[1] * A0[d + j];
            e0 += pv[0] * A0[32] + pv[1] * A0[d + 32];
            e1 += pv[0] * A0[33] + pv[1] * A0[d + 33];
            if (d < 34) e1 = 0.f;
            if (d < 33) e0 = 0.f;
            warp_store_rows34(stage, v, e0, e1, lane, a.heads + grow0 * (NH * d) + h * d, NH * d, nrows, d);
            // operand of the fc GEMM: K index h * 48 + j -> 16-byte units 6h .. 6h+5 of the head-padded row
            float tail[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) tail[j] = 0.f;
            tail[0] = e0; tail[1] = e1;
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int g = 6 * h + u;
                uint8_t* hi = R + (2 * (g >> 3)) * TILE_BYTES;
                store_unit(hi, hi + TILE_BYTES, row, g & 7, u < 4 ? &v[8 * u] : &tail[8 * (u - 4)]);
            }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->hand[1]);
        if (wt == 0) trace(15);

        mbar_wait(&bar->dfull[1], 0);
        tc_fence_after();
        if (wt == 0) trace(16);
        const unsigned long long seed1 = a.seed1 + ((a.p1 > 0.f && a.seed_off) ? *a.seed_off : 0ull);
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = hf * 64 + cc * 32;
            float r[32];
            warp_load_block(stage, r, lane, a.x + grow0 * DM + c0, DM, nrows);
            tmem_ld32(tlane + DB_COL + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bfc[c0 + j];
            if (a.p1 > 0.f) dropout32(v, a.p1, seed1, (unsigned long long)grow * DM + c0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(a.sign, v[j], r[j]);
            warp_store_block(stage, v, lane, a.x1 + grow0 * DM + c0, DM, nrows);
        }
        tc_fence_before();
        if (wt == 0) trace(17);
    }
    tmem_free_512(tmem, warp);
    if (threadIdx.x == 0) trace(18);
}

// ---------------------------------------------------------------------------------------------------------------
// kernel 3: feed-forward block with the hidden tile kept on chip
// ---------------------------------------------------------------------------------------------------------------
struct FfnArgs {
    const uint8_t* w1tiles;     // 8 tiles: (N block 0/1, K chunk 0/1, hi/lo)
    const uint8_t* w2tiles;     // 8 tiles: (K chunk 0..3, hi/lo)
    const float* vec;
    const float* x1;            // (B n, 128) fp32 (residual)
    float* hbuf;                // (B n, 256): drop(relu(.)) -- saved for backward
    float* x2;
    float pf, p2, rscale;
    unsigned long long seedf, seed2;
    const unsigned long long* seed_off;
    int B, n, tiles;
};
constexpr int FFN_RING = 3;
constexpr int FFN_SMEM = 8 * TILE_BYTES + FFN_RING * TILE_BYTES + NWORK * STAGE_BYTES + (DFF + DM) * 4 + 512 + 1024;

__global__ void __launch_bounds__(THREADS, 1) enc_ffn_kernel(const __grid_constant__ CUtensorMap mapX, FfnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_base(smem_raw);
    uint8_t* R = smem;                                         // x1 tile (first 4 tiles) -> hidden operand (8 tiles)
    uint8_t* ring = R + 8 * TILE_BYTES;
    float* staging = reinterpret_cast<float*>(ring + FFN_RING * TILE_BYTES);
    float* b1 = staging + NWORK * 32 * STAGE_PITCH;            // [256]
    float* b2 = b1 + DFF;                                      // [128]
    Bars* bar = reinterpret_cast<Bars*>(b2 + DM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x % a.tiles;
    const int t0 = tile * TM;
    const int nvalid = min(TM, a.n - t0);

    if (threadIdx.x == 0) {
        mbar_init(&bar->xfull, 1);
        mbar_init(&bar->xconv, NWORK);
        for (int s = 0; s < FFN_RING; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar->dfull[i], 1); mbar_init(&bar->hand[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    }
    if (threadIdx.x == 0) { trace_entry(); trace(0); }
    const uint32_t tmem = tmem_alloc_512(&bar->tmem_slot, warp);
    grid_dep_sync();
    if (threadIdx.x == 0) trace(1);
    constexpr int DY_COL = 256;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&bar->xfull, 4 * TILE_BYTES);
            for (int j = 0; j < 4; ++j) tma_load_3d(R + j * TILE_BYTES, &mapX, &bar->xfull, 32 * j, t0, b);
            for (int i = 0; i < 16; ++i) {
                const int s = i % FFN_RING;
                mbar_wait(&bar->empty[s], ((i / FFN_RING) & 1) ^ 1);
                mbar_expect_tx(&bar->full[s], TILE_BYTES);
                const uint8_t* src = i < 8 ? a.w1tiles + (size_t)i * TILE_BYTES : a.w2tiles + (size_t)(i - 8) * TILE_BYTES;
                bulk_load(ring + s * TILE_BYTES, src, TILE_BYTES, &bar->full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ra = smem_u32(R), rb = smem_u32(ring);
            const uint32_t id128 = idesc_bf16(128);
            mbar_wait(&bar->xconv, 0);
            tc_fence_after();
            trace(4);
            for (int i = 0; i < 8; ++i) {                      // hidden = x1 W1^T : two N blocks of 128
                const int s = i % FFN_RING;
                mbar_wait(&bar->full[s], (i / FFN_RING) & 1);
                tc_fence_after();
                const int nb = i >> 2, kc = (i >> 1) & 1;
                mma_weight_tile(tmem + nb * 128, ra + (2 * kc) * TILE_BYTES, ra + (2 * kc + 1) * TILE_BYTES,
                                rb + s * TILE_BYTES, (i & 1) == 0, id128, kc == 0);
                tc_commit(&bar->empty[s]);
                if ((i & 3) == 3) tc_commit(&bar->dfull[nb]);
            }
            trace(5);
            for (int i = 8; i < 16; ++i) {                     // y = hidden W2^T : K chunks follow the epilogue
                const int s = i % FFN_RING;
                const int kc = (i - 8) >> 1;
                if (((i - 8) & 1) == 0) {
                    mbar_wait(&bar->hand[kc], 0);
                    tc_fence_after();
                    trace(20 + kc);
                }
                mbar_wait(&bar->full[s], (i / FFN_RING) & 1);
                tc_fence_after();
                mma_weight_tile(tmem + DY_COL, ra + (2 * kc) * TILE_BYTES, ra + (2 * kc + 1) * TILE_BYTES,
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
        const long long grow = (long long)b * a.n + t0 + row;
        const long long grow0 = (long long)b * a.n + t0 + q * 32;
        const int nrows = max(0, min(32, nvalid - q * 32));
        float* stage = staging + w * 32 * STAGE_PITCH;
        b1[wt] = a.vec[VEC_B1 + wt];
        if (wt < DM) b2[wt] = a.vec[VEC_B2 + wt];
        mbar_wait(&bar->xfull, 0);
        if (wt == 0) trace(9);
        split_tile_inplace(R, wt & 127, wt >> 7);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->xconv);
        worker_bar();
        if (wt == 0) trace(10);

        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const unsigned long long so = a.seed_off ? *a.seed_off : 0ull;
        float v[32];
        // hidden columns hf*128 .. +127 (N block hf): bias, ReLU, dropout; saved to HBM and handed to the second GEMM
        mbar_wait(&bar->dfull[hf], 0);
        tc_fence_after();
        if (wt == 0) trace(11);
        if (wt == 128) trace(12);
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = hf * 128 + cc * 32;
            tmem_ld32(tlane + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + b1[c0 + j], 0.f);
            if (a.pf > 0.f) dropout32(v, a.pf, a.seedf + so, (unsigned long long)grow * DFF + c0);
            warp_store_block(stage, v, lane, a.hbuf + grow0 * DFF + c0, DFF, nrows);
            if (cc == 0 && hf == 0) {                          // chunks 0,1 alias the x1 operand: wait for block 1's MMAs
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
        // y columns hf*64 .. +63
        if (wt == 0) trace(13);
        if (wt == 128) trace(14);
        mbar_wait(&bar->dfull[2], 0);
        tc_fence_after();
        if (wt == 0) trace(15);
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = hf * 64 + cc * 32;
            float r[32];
            warp_load_block(stage, r, lane, a.x1 + grow0 * DM + c0, DM, nrows);
            tmem_ld32(tlane + DY_COL + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += b2[c0 + j];
            if (a.p2 > 0.f) dropout32(v, a.p2, a.seed2 + so, (unsigned long long)grow * DM + c0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(a.rscale, v[j], r[j]);
            warp_store_block(stage, v, lane, a.x2 + grow0 * DM + c0, DM, nrows);
        }
        tc_fence_before();
        if (wt == 0) trace(16);
    }
    tmem_free_512(tmem, warp);
    if (threadIdx.x == 0) trace(17);
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: fp32 parameters -> bf16 hi/lo tile images in streaming order + one fp32 vector block
// ---------------------------------------------------------------------------------------------------------------
struct TileDesc {
    const float* src;
    int ld, r0, c0, rlim, clim, mode, d;
};
// mode 0: T[r][k] = src[(r0+r) * ld + (c0+k)]                         rows < rlim, cols < clim
// mode 1: T[r][k] = src[(c0+k) * ld + (r0+r)]            (transposed) src cols (r0+r) < rlim, src rows (c0+k) < clim
// mode 2: T[r][k] = src[(r0+r) * ld + h*d + j], (h, j) = divmod(c0+k, 48), j < d              (head-padded columns)
// mode 3: T[r][k] = src[(c0+k) * ld + h*d + j], (h, j) = divmod(r0+r, 48), j < d   (head-padded rows, transposed)
constexpr int PACK_MAX_PAIRS = 40;
struct PackTable {
    TileDesc t[PACK_MAX_PAIRS];
    const float* vsrc[32];
    int voff[33];
    int nvec;
};

__global__ void __launch_bounds__(256) enc_pack_kernel(const __grid_constant__ PackTable tab, uint8_t* __restrict__ tiles,
                                                       float* __restrict__ vec, int npairs) {
    if ((int)blockIdx.x >= npairs) {       // vector block
        const int seg = blockIdx.x - npairs;
        const float* s = tab.vsrc[seg];
        const int n = tab.voff[seg + 1] - tab.voff[seg];
        for (int i = threadIdx.x; i < n; i += blockDim.x) vec[tab.voff[seg] + i] = s ? s[i] : 0.f;
        return;
    }
    const TileDesc& t = tab.t[blockIdx.x];
    uint8_t* hi = tiles + (size_t)blockIdx.x * 2 * TILE_BYTES;
    uint8_t* lo = hi + TILE_BYTES;
    for (int e = threadIdx.x; e < 128 * 8; e += blockDim.x) {
        // consecutive threads take consecutive rows of one unit for the transposed modes (coalesced source reads)
        int r, u;
        if (t.mode == 1 || t.mode == 3) { r = e & 127; u = e >> 7; }
        else { r = e >> 3; u = e & 7; }
        float x8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = u * 8 + i;
            float val = 0.f;
            if (t.mode == 0) {
                if (t.r0 + r < t.rlim && t.c0 + k < t.clim) val = t.src[(long long)(t.r0 + r) * t.ld + t.c0 + k];
            } else if (t.mode == 1) {
                if (t.r0 + r < t.rlim && t.c0 + k < t.clim) val = t.src[(long long)(t.c0 + k) * t.ld + t.r0 + r];
            } else if (t.mode == 2) {
                const int kk = t.c0 + k, h = kk / HP, j = kk % HP;
                if (t.r0 + r < t.rlim && kk < t.clim && j < t.d) val = t.src[(long long)(t.r0 + r) * t.ld + h * t.d + j];
            } else {
                const int rr = t.r0 + r, h = rr / HP, j = rr % HP;
                if (rr < t.rlim && t.c0 + k < t.clim && j < t.d) val = t.src[(long long)(t.c0 + k) * t.ld + h * t.d + j];
            }
            x8[i] = val;
        }
        store_unit(hi, lo, r, u, x8);
    }
}

}  // namespace enc
}  // namespace gb200

using namespace gb200;
using namespace gb200::enc;

extern "C" int gb200_encoder_set_trace(unsigned long long* device_buffer) {
    return cudaMemcpyToSymbol(g_enc_trace, &device_buffer, sizeof(device_buffer)) == cudaSuccess ? GB200_OK : GB200_ERR_CUDA;
}

extern "C" int gb200_encoder_supported(int d_model, int n_head, int pos_dim, int d_ff) {
    return d_model == DM && n_head == NH && d_ff == DFF && pos_dim >= 1 && pos_dim <= 2 && encode_fn() != nullptr;
}

extern "C" size_t gb200_encoder_pack_bytes(int d_model, int n_head, int pos_dim, int d_ff) {
    if (!(d_model == DM && n_head == NH && d_ff == DFF)) return 0;
    return (size_t)GB200_ENC_TILES * TILE_BYTES + (size_t)VEC_FLOATS * 4;
}

extern "C" int gb200_encoder_pack(int device, const gb200_encoder_params* P, void* packed, void* stream) {
    use_device(device);
    GB_REQUIRE(P && packed, "gb200_encoder_pack: null argument");
    GB_REQUIRE(gb200_encoder_supported(P->d_model, P->n_head, P->pos_dim, P->d_ff),
               "gb200_encoder_pack: unsupported layer shape (d_model %d, heads %d, pos_dim %d, d_ff %d)", P->d_model,
               P->n_head, P->pos_dim, P->d_ff);
    GB_REQUIRE(((uintptr_t)packed % 128) == 0, "gb200_encoder_pack: packed buffer must be 128-byte aligned");
    const int d = DK + P->pos_dim;
    PackTable tab;
    memset(&tab, 0, sizeof(tab));
    int np = 0;
    auto add = [&](const float* src, int ld, int r0, int c0, int rlim, int clim, int mode) {
        TileDesc& t = tab.t[np++];
        t.src = src; t.ld = ld; t.r0 = r0; t.c0 = c0; t.rlim = rlim; t.clim = clim; t.mode = mode; t.d = d;
    };
    const float* wqkv[3] = {P->wq, P->wk, P->wv};
    // ---- forward stream ----
    for (int nb = 0; nb < 3; ++nb)
        for (int kc = 0; kc < 2; ++kc) add(wqkv[nb], DM, 0, kc * 64, DM, DM, 0);
    for (int kc = 0; kc < 3; ++kc) add(P->wfc, NH * d, 0, kc * 64, DM, NH * HP, 2);
    for (int nb = 0; nb < 2; ++nb)
        for (int kc = 0; kc < 2; ++kc) add(P->w1, DM, nb * 128, kc * 64, DFF, DM, 0);
    for (int kc = 0; kc < 4; ++kc) add(P->w2, DFF, 0, kc * 64, DM, DFF, 0);
    // ---- backward stream (input-gradient GEMMs read the transposed weights) ----
    //   g1 = g2 W2          : B[n = hidden 256][k = out 128] = W2^T          N blocks 2 x K chunks 2
    for (int nb = 0; nb < 2; ++nb)
        for (int kc = 0; kc < 2; ++kc) add(P->w2, DFF, nb * 128, kc * 64, DFF, DM, 1);
    //   dx1 = g1 W1         : B[n = in 128][k = hidden 256] = W1^T           K chunks 4
    for (int kc = 0; kc < 4; ++kc) add(P->w1, DM, 0, kc * 64, DM, DFF, 1);
    //   dheads = dx1 W_fc   : B[n = h*48+j (192)][k = out 128] = W_fc'^T     N blocks 2 x K chunks 2
    for (int nb = 0; nb < 2; ++nb)
        for (int kc = 0; kc < 2; ++kc) add(P->wfc, NH * d, nb * 128, kc * 64, NH * HP, DM, 3);
    //   dx = dqkv W_qkv     : B[n = in 128][k = 384] = W_qkv^T               K chunks 6
    for (int kc = 0; kc < 6; ++kc) add(wqkv[kc / 2], DM, 0, (kc & 1) * 64, DM, DM, 1);
    GB_REQUIRE(np * 2 == GB200_ENC_TILES, "gb200_encoder_pack: internal tile count %d", np * 2);
    int nv = 0, off = 0;
    auto addv = [&](const float* s, int n) { tab.vsrc[nv] = s; tab.voff[nv] = off; off += n; ++nv; };
    addv(P->bq, DM); addv(P->bk, DM); addv(P->bv, DM);
    for (int h = 0; h < NH; ++h) addv(P->gamma_k[h], DK);
    for (int h = 0; h < NH; ++h) addv(P->beta_k[h], DK);
    for (int h = 0; h < NH; ++h) addv(P->gamma_v[h], DK);
    for (int h = 0; h < NH; ++h) addv(P->beta_v[h], DK);
    addv(P->bfc, DM); addv(P->b1, DFF); addv(P->b2, DM);
    tab.voff[nv] = off;
    tab.nvec = nv;
    GB_REQUIRE(off == VEC_FLOATS, "gb200_encoder_pack: internal vector size %d", off);
    uint8_t* tiles = reinterpret_cast<uint8_t*>(packed);
    float* vec = reinterpret_cast<float*>(tiles + (size_t)GB200_ENC_TILES * TILE_BYTES);
    enc_pack_kernel<<<np + nv, 256, 0, as_stream(stream)>>>(tab, tiles, vec, np);
    return check_launch("gb200_encoder_pack");
}

extern "C" size_t gb200_encoder_workspace_bytes(int B, int n, int n_head, int d_k, int pos_dim) {
    const int tiles = (n + TM - 1) / TM, d = d_k + pos_dim;
    return (size_t)B * tiles * n_head * d * d * sizeof(float);
}

extern "C" int gb200_encoder_layer_fwd(int device, const void* packed, int d_model, int n_head, int pos_dim, int d_ff,
                                       const float* x, const float* pos, int B, int n, int has_norm, float eps,
                                       float attn_scale, const unsigned char* keep_mask, float mask_p,
                                       unsigned long long mask_seed, float p_attn_out, unsigned long long seed_attn_out,
                                       float res_sign, float p_ffn, unsigned long long seed_ffn, float p_out,
                                       unsigned long long seed_out, float* qkv, float* rstd_k, float* rstd_v, float* attn,
                                       float* heads, float* x1, float* hidden, float* x2, float* workspace,
                                       size_t workspace_bytes, int stages, void* stream) {
    use_device(device);
    GB_REQUIRE(gb200_encoder_supported(d_model, n_head, pos_dim, d_ff), "gb200_encoder_layer_fwd: unsupported layer shape");
    GB_REQUIRE(packed && x && qkv && attn && heads && x1 && hidden && x2 && workspace, "gb200_encoder_layer_fwd: null buffer");
    GB_REQUIRE(pos_dim == 0 || pos, "gb200_encoder_layer_fwd: pos is null");
    GB_REQUIRE(!has_norm || (rstd_k && rstd_v), "gb200_encoder_layer_fwd: null rstd");
    GB_REQUIRE(B >= 1 && n >= 1, "gb200_encoder_layer_fwd: empty batch");
    GB_REQUIRE(workspace_bytes >= gb200_encoder_workspace_bytes(B, n, n_head, DK, pos_dim),
               "gb200_encoder_layer_fwd: workspace too small");
    GB_REQUIRE(mask_p >= 0.f && mask_p < 1.f && p_attn_out >= 0.f && p_attn_out < 1.f && p_ffn >= 0.f && p_ffn < 1.f &&
                   p_out >= 0.f && p_out < 1.f, "gb200_encoder_layer_fwd: dropout probability outside [0,1)");
    auto al16 = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
    GB_REQUIRE(al16(x) && al16(qkv) && al16(x1) && al16(hidden) && al16(x2) && al16(heads),
               "gb200_encoder_layer_fwd: buffers must be 16-byte aligned");
    static bool configured = false;
    if (!configured) {
        set_smem(enc_qkv_kernel, QKV_SMEM);
        set_smem(enc_attn_kernel, ATT_SMEM);
        set_smem(enc_ffn_kernel, FFN_SMEM);
        configured = true;
    }
    const int tiles = (n + TM - 1) / TM;
    const uint8_t* wt = reinterpret_cast<const uint8_t*>(packed);
    const float* vec = reinterpret_cast<const float*>(wt + (size_t)GB200_ENC_TILES * TILE_BYTES);
    cudaStream_t st = as_stream(stream);
    int launched = 0;
    if (stages & 1) {
        CUtensorMap mx;
        GB_REQUIRE(make_tile_map(&mx, x, DM, n, B), "gb200_encoder_layer_fwd: tensor map (x) failed");
        QkvArgs a;
        a.wtiles = wt + (size_t)TS_QKV * TILE_BYTES; a.vec = vec; a.pos = pos; a.qkv = qkv; a.rstd_k = rstd_k;
        a.rstd_v = rstd_v; a.part = workspace; a.B = B; a.n = n; a.p = pos_dim; a.tiles = tiles; a.has_norm = has_norm;
        a.eps = eps;
        launch_enc(enc_qkv_kernel, B * tiles, QKV_SMEM, st, mx, a);
        ++launched;
    }
    if (stages & 2) {
        CUtensorMap mq;
        GB_REQUIRE(make_tile_map(&mq, qkv, 3 * DM, n, B), "gb200_encoder_layer_fwd: tensor map (qkv) failed");
        AttnArgs a;
        a.wtiles = wt + (size_t)TS_FC * TILE_BYTES; a.vec = vec; a.pos = pos; a.part = workspace; a.keep_mask = keep_mask;
        a.mask_p = keep_mask ? 0.f : mask_p; a.mask_seed = mask_seed; a.seed_off = rng_offset_ptr(); a.scale = attn_scale;
        a.attn = attn; a.heads = heads; a.x = x; a.x1 = x1; a.p1 = p_attn_out; a.seed1 = seed_attn_out; a.sign = res_sign;
        a.B = B; a.n = n; a.p = pos_dim; a.tiles = tiles;
        launch_enc(enc_attn_kernel, B * tiles, ATT_SMEM, st, mq, a);
        ++launched;
    }
    if (stages & 4) {
        CUtensorMap m1;
        GB_REQUIRE(make_tile_map(&m1, x1, DM, n, B), "gb200_encoder_layer_fwd: tensor map (x1) failed");
        FfnArgs a;
        a.w1tiles = wt + (size_t)TS_W1 * TILE_BYTES; a.w2tiles = wt + (size_t)TS_W2 * TILE_BYTES; a.vec = vec; a.x1 = x1;
        a.hbuf = hidden; a.x2 = x2; a.pf = p_ffn; a.p2 = p_out; a.rscale = 1.f; a.seedf = seed_ffn; a.seed2 = seed_out;
        a.seed_off = rng_offset_ptr(); a.B = B; a.n = n; a.tiles = tiles;
        launch_enc(enc_ffn_kernel, B * tiles, FFN_SMEM, st, m1, a);
        ++launched;
    }
    return check_launch("gb200_encoder_layer_fwd", launched);
}
