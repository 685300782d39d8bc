// Micro-benchmark behind DESIGN.md's "why config 4 is not a cluster / DSMEM kernel": per-SM bandwidth of remote
// shared-memory writes (the transposition a one-plane-per-cluster 2-D FFT needs) against local shared-memory writes and
// against an L2-resident global round trip of the same volume.   nvcc -O3 -arch=sm_100a -o dsmem_bw dsmem_bw.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int kBytes = 96 * 1024;   // per CTA buffer
constexpr int kIters = 64;

template <int MODE>   // 0: local smem, 1: remote smem of the next CTA of the cluster (transposition pattern: CTA r -> r+1+k)
__global__ void __launch_bounds__(512) smem_write_kernel(float* sink, long long* cycles) {
    extern __shared__ __align__(16) unsigned char buf[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), nr = cluster.num_blocks();
    float4* local = reinterpret_cast<float4*>(buf);
    const int n4 = kBytes / 16;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) local[i] = make_float4(i, 0, 0, 0);
    cluster.sync();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        float4* dst = MODE == 0 ? local : reinterpret_cast<float4*>(cluster.map_shared_rank(buf, (rank + 1 + it % (nr - 1)) % nr));
        const float4 v = make_float4(it, threadIdx.x, 0, 0);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = v;
    }
    cluster.sync();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (threadIdx.x == 0 && local[1].x == -1.f) sink[0] = local[2].y;
}

__global__ void __launch_bounds__(512) l2_roundtrip_kernel(float4* scratch, float* sink, long long* cycles) {
    // every CTA writes kBytes to an L2-resident scratch slot and reads the slot of CTA + 1 back (bulk of an L2 transposition)
    const int n4 = kBytes / 16;
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        float4* mine = scratch + (size_t)blockIdx.x * n4;
        const float4* other = scratch + (size_t)((blockIdx.x + 1 + it) % gridDim.x) * n4;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) mine[i] = make_float4(it, i, 0, 0);
        __syncthreads();
        for (int i = threadIdx.x; i < n4; i += blockDim.x) { const float4 q = __ldcg(other + i); acc.x += q.x; acc.y += q.y; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc.x == -1.f) sink[0] = acc.y;
}

int main() {
    float* sink; long long* cyc; float4* scratch;
    cudaMalloc(&sink, 16); cudaMalloc(&cyc, 1024 * sizeof(long long));
    long long h[1024];
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int mode = 0; mode < 2; ++mode) {
        for (int csz : {2, 4, 8}) {
            auto kern = mode == 0 ? smem_write_kernel<0> : smem_write_kernel<1>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes);
            cudaLaunchConfig_t cfg = {};
            const int grid = (sms / csz) * csz;
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = kBytes;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)csz, 1, 1};
            cfg.attrs = at; cfg.numAttrs = 1;
            int maxc = 0; cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg);
            for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, kern, sink, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
            printf("%s cluster=%d grid=%d (max active clusters %d): %.1f B/clk/SM  [%s]\n", mode ? "remote smem write" : "local  smem write", csz, grid, maxc,
                   (double)kBytes * kIters / avg, cudaGetErrorString(e));
        }
    }
    cudaMalloc(&scratch, (size_t)sms * kBytes);
    for (int rep = 0; rep < 2; ++rep) l2_roundtrip_kernel<<<sms, 512>>>(scratch, sink, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("L2 round trip (write %d KB + read %d KB per CTA and iteration): %.1f B/clk/SM written and as much read  [%s]\n", kBytes / 1024, kBytes / 1024,
           (double)kBytes * kIters / avg, cudaGetErrorString(e));
    return 0;
}
