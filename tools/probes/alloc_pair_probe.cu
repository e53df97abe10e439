// Can two CTA pairs that share an SM pair deadlock in tcgen05.alloc.cta_group::2?
// Each CTA (cluster of 2, two CTAs resident per SM via 100 KB of dynamic smem) spins a pseudo-random number of cycles, then its
// warp 0 allocates 256 TMEM columns as a pair, relinquishes, cluster-syncs and frees.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o alloc_pair_probe alloc_pair_probe.cu
// Run: ./alloc_pair_probe <launches> <jitter_cycles>; a hang (use `timeout`) answers yes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../mapf_gpt_b200/csrc/ptx.cuh"
using namespace mg;

__global__ void __launch_bounds__(128, 2) probe(int jitter, unsigned seed, unsigned long long *sink, unsigned *smid_of)
{
    if (threadIdx.x == 0) {   // where did the hardware put rank 0 / rank 1 of this cluster?
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        smid_of[blockIdx.x] = smid;
    }
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *slot = reinterpret_cast<unsigned *>(smem);
    const int warp = threadIdx.x >> 5;
    unsigned h = (blockIdx.x * 2654435761u) ^ seed;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    if (jitter > 0) {
        const long long t0 = clock64(), d = h % (unsigned)jitter;
        while (clock64() - t0 < d) { }
    }
    if (warp == 0) tmem_alloc_pair<256>(slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const unsigned t = *slot;
    if (threadIdx.x == 0 && t == 0xffffffffu) atomicAdd(sink, 1ull);
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair<256>(t);
}

int main(int argc, char **argv)
{
    const int launches = argc > 1 ? atoi(argv[1]) : 20, jitter = argc > 2 ? atoi(argv[2]) : 2000;
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    const int nblk = 148 * 2 * 64;
    unsigned *smid_of;
    cudaMalloc(&smid_of, nblk * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int i = 0; i < launches; i++) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(148 * 2 * 64);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 100 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t rc = cudaLaunchKernelEx(&cfg, probe, jitter, (unsigned)i * 7919u + 1u, sink, smid_of);
        if (rc != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(rc)); return 1; }
        rc = cudaDeviceSynchronize();
        if (rc != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(rc)); return 1; }
    }
    {   // orientation statistics of the last launch
        unsigned *h = (unsigned *)malloc(nblk * 4);
        cudaMemcpy(h, smid_of, nblk * 4, cudaMemcpyDeviceToHost);
        int same_tpc = 0, leader_even = 0, leader_low = 0;
        for (int p = 0; p < nblk / 2; p++) {
            const unsigned a = h[2 * p], b = h[2 * p + 1];
            same_tpc += (a / 2 == b / 2);
            leader_even += (a % 2 == 0);
            leader_low += (a < b);
        }
        printf("pairs %d: both ranks in SM pair (2k,2k+1): %d, rank 0 on the even SM: %d, rank 0 on the lower SM id: %d\n", nblk / 2, same_tpc,
               leader_even, leader_low);
    }
    printf("%d launches of %d CTA pairs completed (jitter %d cycles)\n", launches, 148 * 64, jitter);
    return 0;
}
