// Microbenchmark (not product code): cta_group::2 tcgen05.mma (M = 256 over an SM pair) from shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_bench umma2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ bool mbar_wait_to(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) return false;
    }
    return true;
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

struct V { int n; int iters; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench2(V v, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("fence.proxy.async.shared::cta;");
    cluster.sync();
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    cluster.sync();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base_s;
    bool ok = true;
    if (rank == 0 && warp == 1 && lane == 0) {
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
        const int nh = v.n / 2;                            // B rows held by each CTA
        uint64_t ad[8], bd[8];
        for (int ks = 0; ks < 8; ++ks) {
            ad[ks] = desc(a_base + ks * 2 * 16 * 128, 16 * 128, 128);
            bd[ks] = desc(b_base + ks * 2 * (nh / 8) * 128, (nh / 8) * 128, 128);
        }
        const uint32_t idesc = (1u << 4) | ((uint32_t)(v.n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        long long t0 = clock64();
        for (int it = 0; it < v.iters; ++it) {
            #pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma2(tb + ((ks & 1) ? (uint32_t)v.n : 0u) * (v.n <= 256 ? 1 : 0) * 0, ad[ks], bd[ks], idesc, 1);
        }
        tc_commit2(smem_u32(&bar));
        ok = mbar_wait_to(smem_u32(&bar), 0);
        long long t1 = clock64();
        out[(blockIdx.x / 2) * 2] = ok ? (t1 - t0) : -1;
    }
    if (rank == 1 && warp == 1 && lane == 0) {
        ok = mbar_wait_to(smem_u32(&bar), 0);              // multicast commit must also reach the peer's barrier
        out[(blockIdx.x / 2) * 2 + 1] = ok ? 1 : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    cluster.sync();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    int ns[] = {256, 240, 224, 208, 192, 176, 160, 144, 128, 64};
    for (int n : ns) {
        V v{n, 256};
        for (int i = 0; i < 148; ++i) out[i] = 0;
        bench2<<<148, 128, 128 * 1024>>>(v, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d failed: %s\n", n, cudaGetErrorString(e)); return 1; }
        double s = 0; int bad = 0;
        for (int b = 0; b < 74; ++b) { if (out[2 * b] < 0 || out[2 * b + 1] != 1) bad++; s += out[2 * b]; }
        printf("cta_group::2 M=256 N=%3d : %.1f cyc/MMA (ideal %d)  bad=%d\n", n, s / 74 / (256 * 8), n / 2, bad);
    }
    return 0;
}
