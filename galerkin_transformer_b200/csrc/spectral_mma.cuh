// Warp-level TF32 tensor-core variants of the two grid-sized spectral kernels (tf32 precision mode).
// Per row R the truncated transforms are small dense products against a twiddle matrix shared by all rows:
//   ydft :  D[j][c]  = sum_Y  A[j][Y] * x[R][Y][c]          j < 2m  (cos rows, then -sin rows)
//   yidft:  y[Y][o]  = act( sum_k A2[Y][k] * B2[k][o] + bias[o] ),  k over 2m spectral terms + Ci channels
// One warp owns one row (ydft) or one row x chunk of 16-row tiles (yidft); the row's operand is read from
// global memory directly in mma fragment order (every element fetched exactly once, 32-byte sectors), the
// twiddle matrix lives in shared memory, accumulators in registers (mma.sync.m16n8k8, fp32 accumulate).
#pragma once
#include "mma_tf32.cuh"

namespace gb200 {

constexpr int SM_WARPS = 4;

// ------------------------------------------------------------------------------------------------
// ydft: out[R][ky][c] = scale * herm(ky) * sum_Y x[R][Y][c] * exp(-i 2 pi ky Y / n);  2m <= 32, C % 8 == 0, C <= 64
// ------------------------------------------------------------------------------------------------
template <int NT>    // n-tiles of 8 channels
__global__ void __launch_bounds__(SM_WARPS * 32) ydft_mma_kernel(const float* __restrict__ x, long long R, int n,
                                                                 int C, int m, const float2* __restrict__ twY,
                                                                 float scale, int hermitian,
                                                                 float2* __restrict__ out) {
    pdl_enter();
    extern __shared__ float As[];                  // [32][AP]  rows j: cos(ky=j) | -sin(ky=j-m) | 0
    const int KP = (n + 7) / 8 * 8;
    const int AP = KP + ((4 - KP % 32 + 32) % 32 == 0 ? 0 : ((4 - KP % 32 + 32) % 32));   // pitch = 4 (mod 32)
    for (int e = threadIdx.x; e < 32 * AP; e += blockDim.x) {
        const int j = e / AP, Y = e % AP;
        float v = 0.f;
        if (Y < n) {
            if (j < m) v = twY[(long long)j * n + Y].x;
            else if (j < 2 * m) v = -twY[(long long)(j - m) * n + Y].y;
        }
        As[e] = to_tf32(v);
    }
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, tq = lane % 4;
    const long long row = (long long)blockIdx.x * SM_WARPS + warp;
    if (row >= R) return;
    const float* xr = x + row * n * C;
    float acc[2][NT][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[a][c][k] = 0.f;
    float bf[2][NT][2];                            // double-buffered B fragments (the row's x values)
    auto loadB = [&](int buf, int k0) {
#pragma unroll
        for (int c = 0; c < NT; ++c) {
            const int col = 8 * c + g;
            const int y0 = k0 + tq, y1 = k0 + tq + 4;
            bf[buf][c][0] = (y0 < n && col < C) ? to_tf32(xr[(long long)y0 * C + col]) : 0.f;
            bf[buf][c][1] = (y1 < n && col < C) ? to_tf32(xr[(long long)y1 * C + col]) : 0.f;
        }
    };
    loadB(0, 0);
    const int nk = KP / 8;
    for (int ks = 0; ks < nk; ++ks) {
        const int cur = ks & 1;
        if (ks + 1 < nk) loadB(cur ^ 1, 8 * (ks + 1));
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            float af[4];
            af[0] = As[(16 * a + g) * AP + 8 * ks + tq];
            af[1] = As[(16 * a + g + 8) * AP + 8 * ks + tq];
            af[2] = As[(16 * a + g) * AP + 8 * ks + tq + 4];
            af[3] = As[(16 * a + g + 8) * AP + 8 * ks + tq + 4];
#pragma unroll
            for (int c = 0; c < NT; ++c) mma_tf32(acc[a][c], af, bf[cur][c]);
        }
    }
    float* o = reinterpret_cast<float*>(out) + row * m * C * 2;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = 16 * a + g + (k >= 2 ? 8 : 0);
                const int col = 8 * c + 2 * tq + (k & 1);
                if (j >= 2 * m || col >= C) continue;
                const int ky = j < m ? j : j - m;
                const float s = scale * ((hermitian && !(ky == 0 || 2 * ky == n)) ? 2.f : 1.f);
                o[((long long)ky * C + col) * 2 + (j < m ? 0 : 1)] = acc[a][c][k] * s;
            }
}

// ------------------------------------------------------------------------------------------------
// yidft + pointwise residual + bias + activation (see yidft_epi_kernel for the maths).
// K layout: [0, 2m) spectral (cos * Zre rows, then -sin * Zim rows), padded to KS = ceil8(2m); then Ci channels.
// ------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(SM_WARPS * 32) yidft_mma_kernel(
    const float2* __restrict__ Z, long long R, int n, int m, int Co, const float2* __restrict__ twY, float scale,
    int hermitian, const float* __restrict__ x2, int Ci, const float* __restrict__ Wm,
    const float* __restrict__ bias, int act, float* __restrict__ y, float* __restrict__ zout, int tiles_per_warp,
    int chunks_per_row) {
    pdl_enter();
    extern __shared__ float A2[];                  // [MP][SP]: twiddle part of the A operand, rows Y
    const int KS = (2 * m + 7) / 8 * 8;            // spectral K, padded
    const int KC = (Ci + 7) / 8 * 8;               // channel K, padded
    const int MP = (n + 15) / 16 * 16;
    const int SP = KS + ((4 - KS % 32 + 32) % 32);  // pitch = 4 (mod 32)
    for (int e = threadIdx.x; e < MP * SP; e += blockDim.x) {
        const int Y = e / SP, k = e % SP;
        float v = 0.f;
        if (Y < n && k < 2 * m) {
            const int ky = k < m ? k : k - m;
            const float2 t = twY[(long long)ky * n + Y];
            const float s = scale * ((hermitian && !(ky == 0 || 2 * ky == n)) ? 2.f : 1.f);
            v = k < m ? t.x * s : -t.y * s;
        }
        A2[e] = to_tf32(v);
    }
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, tq = lane % 4;
    const long long wid = (long long)blockIdx.x * SM_WARPS + warp;
    const long long row = wid / chunks_per_row;
    const int chunk = (int)(wid % chunks_per_row);
    if (row >= R) return;
    const int nks = KS / 8, nkc = KC / 8;
    // B fragments: spectral rows from Z[row] (re, then im), channel rows from Wm; all kept in registers
    constexpr int MAXKS = 4, MAXKC = 8;            // 2m <= 32, Ci <= 64
    float bs[MAXKS][NT][2], bc[MAXKC][NT][2];
    const float2* zr = Z + row * m * Co;
#pragma unroll
    for (int ks = 0; ks < MAXKS; ++ks)
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 8 * ks + tq + 4 * h, col = 8 * c + g;
                float v = 0.f;
                if (ks < nks && k < 2 * m && col < Co) {
                    const float2 zz = zr[(long long)(k < m ? k : k - m) * Co + col];
                    v = k < m ? zz.x : zz.y;
                }
                bs[ks][c][h] = to_tf32(v);
            }
#pragma unroll
    for (int kc = 0; kc < MAXKC; ++kc)
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 8 * kc + tq + 4 * h, col = 8 * c + g;
                bc[kc][c][h] = (kc < nkc && i < Ci && col < Co) ? to_tf32(Wm[(long long)i * Co + col]) : 0.f;
            }
    const float* xr = x2 + row * n * Ci;
    const int mt0 = chunk * tiles_per_warp, mt1 = min(MP / 16, mt0 + tiles_per_warp);
    for (int mt = mt0; mt < mt1; ++mt) {
        const int y0 = 16 * mt + g, y1 = y0 + 8;
        float acc[NT][4];
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[c][k] = 0.f;
        // channel part first (global loads in flight while the spectral part runs from shared memory)
        float ax[MAXKC][4];
#pragma unroll
        for (int kc = 0; kc < MAXKC; ++kc) {
            const int i0 = 8 * kc + tq, i1 = i0 + 4;
            ax[kc][0] = (kc < nkc && y0 < n && i0 < Ci) ? xr[(long long)y0 * Ci + i0] : 0.f;
            ax[kc][1] = (kc < nkc && y1 < n && i0 < Ci) ? xr[(long long)y1 * Ci + i0] : 0.f;
            ax[kc][2] = (kc < nkc && y0 < n && i1 < Ci) ? xr[(long long)y0 * Ci + i1] : 0.f;
            ax[kc][3] = (kc < nkc && y1 < n && i1 < Ci) ? xr[(long long)y1 * Ci + i1] : 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < MAXKS; ++ks) {
            if (ks < nks) {
                float af[4];
                af[0] = A2[y0 * SP + 8 * ks + tq];
                af[1] = A2[y1 * SP + 8 * ks + tq];
                af[2] = A2[y0 * SP + 8 * ks + tq + 4];
                af[3] = A2[y1 * SP + 8 * ks + tq + 4];
#pragma unroll
                for (int c = 0; c < NT; ++c) mma_tf32(acc[c], af, bs[ks][c]);
            }
        }
#pragma unroll
        for (int kc = 0; kc < MAXKC; ++kc) {
            if (kc < nkc) {
                float af[4] = {to_tf32(ax[kc][0]), to_tf32(ax[kc][1]), to_tf32(ax[kc][2]), to_tf32(ax[kc][3])};
#pragma unroll
                for (int c = 0; c < NT; ++c) mma_tf32(acc[c], af, bc[kc][c]);
            }
        }
#pragma unroll
        for (int c = 0; c < NT; ++c) {
            const int col = 8 * c + 2 * tq;
            if (col >= Co) continue;
            const float b0 = bias ? bias[col] : 0.f;
            const float b1 = (bias && col + 1 < Co) ? bias[col + 1] : 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int Y = h ? y1 : y0;
                if (Y >= n) continue;
                const float z0 = acc[c][2 * h] + b0, z1 = acc[c][2 * h + 1] + b1;
                const long long idx = (row * n + Y) * Co + col;
                if (col + 1 < Co && (Co % 2 == 0)) {
                    if (zout) *reinterpret_cast<float2*>(zout + idx) = make_float2(z0, z1);
                    *reinterpret_cast<float2*>(y + idx) = make_float2(act_apply(act, z0), act_apply(act, z1));
                } else {
                    if (zout) zout[idx] = z0;
                    y[idx] = act_apply(act, z0);
                    if (col + 1 < Co) {
                        if (zout) zout[idx + 1] = z1;
                        y[idx + 1] = act_apply(act, z1);
                    }
                }
            }
        }
    }
}

}  // namespace gb200
