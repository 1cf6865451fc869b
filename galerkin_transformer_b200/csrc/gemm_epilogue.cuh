// Epilogue shared by the SIMT and tcgen05 GEMMs:
//   v = alpha*acc + bias[n];  Z = v;  v = act(v);  v *= act'(G) [gate];  v *= dropout(seed, element);
//   C = R + rscale*v  (+= C)
// The gate turns a GEMM into "input gradient of the NEXT layer times the activation/dropout derivative of THIS
// layer" (g1 = (g2 W2) * act'(z1) * mask1), i.e. the elementwise backward pass between two Linear layers fused
// into the producing GEMM.
#pragma once
#include "common.cuh"

namespace gb200 {

struct GemmEpilogue {
    float* C; int ldc; long long sC;
    float alpha;
    const float* bias;
    int act;
    float* Z; int ldz;
    float drop_p; unsigned long long seed;
    const unsigned long long* seed_off;   // device-side step counter (or null): seed += *seed_off
    const float* R; int ldr; float rscale;
    int accumulate;
    const float* G; int ldg; int gate;     // gate: ACT_RELU -> keep where G > 0; ACT_SILU -> times silu'(G); null = off
    // fused per-head LayerNorm statistics (tcgen05 float4 epilogue only): columns [hn_lo, hn_hi) are split into
    // groups of hn_dk; each group of a row is replaced by (v - mean) * rstd and rstd goes to hn_rstd[block][row, head]
    int hn_dk, hn_lo, hn_hi, hn_heads;
    float hn_eps;
    float* hn_rstd[2];

    __device__ __forceinline__ void store(int batch, int m, int n, int M, int N, float acc) const {
        float v = alpha * acc;
        if (bias) v += bias[n];
        if (Z) Z[(long long)batch * sC + (long long)m * ldz + n] = v;
        v = act_apply(act, v);
        if (G) {
            const float gv = G[(long long)batch * sC + (long long)m * ldg + n];
            v = gate == ACT_RELU ? (gv > 0.f ? v : 0.f) : v * act_grad(gate, gv);
        }
        if (drop_p > 0.f)
            v *= dropout_scale(drop_p, seed + (seed_off ? *seed_off : 0ull), ((unsigned long long)batch * M + m) * N + n);
        float* c = C + (long long)batch * sC + (long long)m * ldc + n;
        if (R) v = R[(long long)batch * sC + (long long)m * ldr + n] + rscale * v;
        else v *= rscale;
        if (accumulate) v += *c;
        *c = v;
    }
};

// gate of the NEXT gb200_gemm / gb200_gemm_tc call on this thread (set by the *_gated entry points)
struct GemmGate { const float* G; int ldg; int act; };
GemmGate& next_gemm_gate();

}  // namespace gb200
