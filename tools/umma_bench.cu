// Microbenchmark (not product code): tcgen05.mma issue rate from shared memory for several operand layouts
// and shapes, and tcgen05.ld throughput.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, int kind) {
    if (kind == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* u) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                   "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                 : "r"(taddr) : "memory");
}

struct Variant { int layout; int n; int kind; int a_tmem; int ld_warps; int mma_iters; int m; int nacc; };

// layout: 0 = no-swizzle, K-adjacent cores far apart (LBO = rows/8*128, SBO = 128)
//         1 = no-swizzle, K-adjacent cores contiguous (LBO = 128, SBO = KD/8*128)
//         2 = SWIZZLE_128B (SBO = 1024, atoms of 64 halfs along K)
__global__ void __launch_bounds__(384, 1) bench(Variant v, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base_s;
    const int KD = 128, ksteps = (v.kind == 0) ? 8 : 4;     // i8: K = 32 per MMA
    long long t0 = 0, t1 = 0;
    __syncthreads();
    if (warp == 1 && lane == 0 && v.mma_iters > 0) {
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
        const int M = v.m;
        uint32_t idesc = (v.kind == 0 ? (1u << 4) : (2u << 4) | (1u << 7) | (1u << 10)) | ((uint32_t)(v.n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint64_t ad[8], bd[8];
        for (int ks = 0; ks < 8; ++ks) {
            if (v.layout == 0) {
                ad[ks] = desc(a_base + ks * 2 * (M / 8) * 128, (M / 8) * 128, 128, 0);
                bd[ks] = desc(b_base + ks * 2 * (v.n / 8) * 128, (v.n / 8) * 128, 128, 0);
            } else if (v.layout == 1) {
                ad[ks] = desc(a_base + ks * 256, 128, (KD / 8) * 128, 0);
                bd[ks] = desc(b_base + ks * 256, 128, (KD / 8) * 128, 0);
            } else {
                ad[ks] = desc(a_base + (ks / 4) * M * 128 + (ks % 4) * 32, 16, 1024, 2);
                bd[ks] = desc(b_base + (ks / 4) * v.n * 128 + (ks % 4) * 32, 16, 1024, 2);
            }
        }
        const uint32_t d0 = tb, d1 = tb + (v.nacc > 1 ? v.n : 0);
        const int a_tmem = v.a_tmem, kind = v.kind;
        t0 = clock64();
        for (int it = 0; it < v.mma_iters; ++it) {
            #pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                if (ks >= ksteps) break;
                const uint32_t dd = (ks & 1) ? d1 : d0;
                if (a_tmem) umma_ts(dd, tb + 448 + ks * 8, bd[ks], idesc, 1);
                else umma(dd, ad[ks], bd[ks], idesc, 1, kind);
            }
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        out[blockIdx.x * 4 + 0] = t1 - t0;
    }
    if (warp >= 4 && warp < 4 + v.ld_warps) {
        uint32_t u[32];
        uint32_t accum = 0;
        const uint32_t ta = tb + ((uint32_t)((warp & 3) * 32) << 16);
        long long s0 = clock64();
        for (int it = 0; it < 256; ++it) {
            tmem_ld32(ta + (it & 7) * 32, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            #pragma unroll
            for (int j = 0; j < 32; ++j) accum += u[j];
        }
        long long s1 = clock64();
        if (lane == 0) { out[blockIdx.x * 4 + 1] = s1 - s0; out[blockIdx.x * 4 + 2] = accum; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 148 * 4 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    Variant vs[] = {
        {0, 128, 0, 0, 0, 256, 128, 1}, {0, 128, 0, 0, 0, 256, 128, 2}, {1, 128, 0, 0, 0, 256, 128, 2}, {2, 128, 0, 0, 0, 256, 128, 2},
        {0, 256, 0, 0, 0, 256, 128, 1}, {0, 256, 0, 0, 0, 256, 128, 2}, {1, 256, 0, 0, 0, 256, 128, 2}, {2, 256, 0, 0, 0, 256, 128, 2},
        {0, 128, 0, 1, 0, 256, 128, 1}, {0, 128, 0, 1, 0, 256, 128, 2}, {0, 256, 0, 1, 0, 256, 128, 1},
        {0, 64, 0, 0, 0, 256, 128, 2}, {0, 192, 0, 0, 0, 256, 128, 2},
        {0, 128, 1, 0, 0, 256, 128, 2}, {0, 256, 1, 0, 0, 256, 128, 2},
        {0, 128, 0, 0, 8, 256, 128, 2}, {0, 256, 0, 0, 8, 256, 128, 2},
    };
    for (auto& v : vs) {
        for (int i = 0; i < 148 * 4; ++i) out[i] = 0;
        bench<<<148, 384, 200 * 1024>>>(v, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant failed: %s\n", cudaGetErrorString(e)); return 1; }
        double mma = 0, ld = 0;
        for (int b = 0; b < 148; ++b) { mma += out[b * 4]; ld += out[b * 4 + 1]; }
        mma /= 148; ld /= 148;
        const int ksteps = v.kind == 0 ? 8 : 4;
        printf("layout=%d N=%3d kind=%s a_tmem=%d ld_warps=%d nacc=%d : ", v.layout, v.n, v.kind ? "i8 " : "f16", v.a_tmem, v.ld_warps, v.nacc);
        if (v.mma_iters) printf("%.1f cyc/MMA (ideal %d)  ", mma / (v.mma_iters * ksteps), v.kind == 0 ? v.n / 2 : v.n / 2);
        if (v.ld_warps) printf("LDTM.x32: %.1f cyc each per warp -> %.1f B/cyc/SM", ld / 256, v.ld_warps * 4096.0 / (ld / 256));
        printf("\n");
    }
    return 0;
}
