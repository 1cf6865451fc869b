// Probe: can 8 clusters of 15 CTAs (one 200 KB CTA per SM) be co-resident on this GPU, and what does a cluster barrier +
// distributed-shared-memory read cost?   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(320, 1) probe(float* out, int words) {
    extern __shared__ float sm[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank(), nb = cl.num_blocks();
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = (float)(rank + 1);
    cl.sync();
    // CTA `rank` sums slice `rank` of every block's buffer through distributed shared memory
    const int per = (words + nb - 1) / nb, e0 = rank * per, e1 = min(words, e0 + per);
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        float s = 0.f;
        for (unsigned k = 0; k < nb; ++k) s += cl.map_shared_rank(sm, k)[e];
        out[(size_t)(blockIdx.x / nb) * words + e] = s;
    }
    cl.sync();
}

int main() {
    const int CL = 15, GRID = 120, SMEM = 200 * 1024, WORDS = 4624;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(GRID); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nclusters = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, probe, &cfg);
    printf("cudaOccupancyMaxActiveClusters(size %d, %d KB): %d (%s)\n", CL, SMEM / 1024, nclusters, cudaGetErrorString(e));
    float* out; cudaMalloc(&out, (size_t)(GRID / CL) * WORDS * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int numattr = 1; numattr <= 2; ++numattr) {
        cfg.numAttrs = numattr;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a);
            e = cudaLaunchKernelEx(&cfg, probe, out, WORDS);
            cudaEventRecord(b);
            cudaError_t s = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, a, b);
            printf("attrs=%d launch: %s / %s   %.1f us\n", numattr, cudaGetErrorString(e), cudaGetErrorString(s), ms * 1e3);
        }
    }
    float h[4]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("out[0] = %.1f (expect %d)\n", h[0], CL * (CL + 1) / 2);
    for (int cl2 = 16; cl2 >= 8; cl2 -= 4) {
        attr[0].val.clusterDim.x = cl2; cfg.numAttrs = 1; cfg.gridDim = dim3(cl2 * 8);
        e = cudaOccupancyMaxActiveClusters(&nclusters, probe, &cfg);
        printf("max active clusters of %d: %d (%s)\n", cl2, nclusters, cudaGetErrorString(e));
    }
    return 0;
}
