// Probe of the tcgen05 shared-memory operand layouts used by tcpair.cu: one kind::tf32 MMA (K = 8) on small integer
// matrices in the canonical NO-SWIZZLE layouts, K-major and MN-major, M = 128 / 64, N = 64 / 16 / 8; prints max |D - A.B|.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcprobe.bin tcprobe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
struct Cfg {
    int M, N, a_mn, b_mn;
    // byte strides of the operand images in shared memory: element (r, k) of A at a_r8 * (r / 8 or r / 4) + ... see fill()
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;     // descriptor fields
    uint32_t a_sk, a_smn, b_sk, b_smn;       // image: byte stride between K groups / between MN groups
    uint32_t a_sw, b_sw;                     // swizzle row bytes (0 = none, 32, 64, 128, 1 = 128B rows with 32B base); descriptor layout type follows
    int es;                                  // element bytes: 4 = tf32 (K = 8), 2 = bf16 (K = 16)
    int rows32;                              // k rows per atom of the 32B-base layout (4 or 8)
    int a_k0;                                // A: first k of the instruction (start address advanced by a_k0 * es; K-major only)
};
__host__ __device__ inline uint32_t layout_of(uint32_t sw) { return sw == 128 ? 2u : sw == 64 ? 4u : sw == 32 ? 6u : sw == 1 ? 1u : 0u; }

// canonical no-swizzle images.  K-major:  (r % 8) * 16 + (r / 8) * Smn + (k % 4) * 4 + (k / 4) * Sk
//                               MN-major: (r % 4) * 4 + (r / 4) * Smn + (k % 8) * 16 + (k / 8) * Sk
// swizzled images (row = sw bytes): K-major: row r % 8, bytes 4 k, groups of 8 r at Smn;  MN-major: row k % 8, bytes 4 (r % (sw / 4)),
// groups of sw / 4 along MN at Smn, groups of 8 k at Sk;  16-byte chunk index ^= (row index bits above 128 B) -- Swizzle<B,4,3> on the byte offset
__device__ uint32_t off(int mn, int r, int k, uint32_t sk, uint32_t smn, uint32_t sw, int es, int rows32) {
    const int T = 16 / es;
    if (sw == 0) return mn ? (r % T) * es + (r / T) * smn + (k % 8) * 16 + (k / 8) * sk : (r % 8) * 16 + (r / 8) * smn + (k % T) * es + (k / T) * sk;
    if (sw == 1 && mn) {  // MN-major: 128-byte rows along MN, rows32 k rows per atom, 32-byte granule index ^= k row & 3
        const int per = 128 / es;
        uint32_t o = (k % rows32) * 128 + (r % per) * es;
        o ^= ((o >> 7) & 3) << 5;
        return o + (r / per) * smn + (k / rows32) * sk;
    }
    if (sw == 1) {        // K-major image of the same thing: row r at r * 128 (8 rows per SBO), k along the row (up to 32), granule ^= r & 3
        uint32_t o = (r % 8) * 128 + k * es;
        o ^= ((o >> 7) & 3) << 5;
        return o + (r / 8) * smn;
    }
    const uint32_t per = sw / es;
    uint32_t o = mn ? (k % 8) * sw + (r % per) * es : (r % 8) * sw + k * es;
    o ^= ((o >> 7) & (sw / 16 - 1)) << 4;
    return o + (mn ? (r / per) * smn + (k / 8) * sk : (r / 8) * smn);
}
__device__ void put(unsigned char *p, float v, int es) {
    if (es == 4) *reinterpret_cast<float *>(p) = v;
    else *reinterpret_cast<unsigned short *>(p) = (unsigned short)(__float_as_uint(v) >> 16);
}
__host__ __device__ float aval(int m, int k) { return (float)((m * 5 + k * 3) % 7 - 3) + 0.5f * (float)(k % 2); }
__host__ __device__ float bval(int n, int k) { return (float)((n * 3 + k) % 5 - 2) + 0.25f * (float)(n % 3); }

__global__ void __launch_bounds__(128, 1) probe(Cfg c, float *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sA = s_u32(smem), sB = sA + 32768;
    for (int i = threadIdx.x; i < 16384; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    __syncthreads();
    const int KK = 32 / c.es;
    for (int i = threadIdx.x; i < c.M * KK; i += 128) {
        const int m = i / KK, k = i % KK;
        put(smem + off(c.a_mn, m, k + c.a_k0, c.a_sk, c.a_smn, c.a_sw, c.es, c.rows32), aval(m, k), c.es);
    }
    for (int i = threadIdx.x; i < c.N * KK; i += 128) {
        const int n = i / KK, k = i % KK;
        put(smem + 32768 + off(c.b_mn, n, k, c.b_sk, c.b_smn, c.b_sw, c.es, c.rows32), bval(n, k), c.es);
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = make_idesc(c.M, c.N, c.a_mn, c.b_mn);
        if (c.es == 2) {
            idesc = (idesc & ~((7u << 7) | (7u << 10))) | (1u << 7) | (1u << 10);   // A = B = BF16
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(make_desc(sA + c.a_k0 * c.es, c.a_lbo, c.a_sbo, layout_of(c.a_sw))), "l"(make_desc(sB, c.b_lbo, c.b_sbo, layout_of(c.b_sw))), "r"(idesc), "r"(0) : "memory");
        } else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(make_desc(sA + c.a_k0 * c.es, c.a_lbo, c.a_sbo, layout_of(c.a_sw))), "l"(make_desc(sB, c.b_lbo, c.b_sbo, layout_of(c.b_sw))), "r"(idesc), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s_u32(&bar)) : "memory");
        if (!done && ++spins > (1 << 22)) __trap();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int col = 0; col < c.N; col += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + col) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 8; ++k) out[(32 * warp + lane) * 64 + col + k] = __uint_as_float(r[k]);   // [tmem lane][column]
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

static void run(const char *name, Cfg c) {
    float *d;
    cudaMalloc(&d, 128 * 64 * 4);
    cudaMemset(d, 0, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<<<1, 128, 65536>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> h(128 * 64);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, mag = 0;
    for (int m = 0; m < c.M; ++m) {
        const int lane = c.M == 64 ? (m % 16) + 32 * (m / 16) : m;
        for (int n = 0; n < c.N; ++n) {
            double ref = 0;
            for (int k = 0; k < 32 / c.es; ++k) ref += (double)aval(m, k) * bval(n, k);
            const double err = fabs(h[lane * 64 + n] - ref);
            if (err > worst) worst = err;
            if (fabs(ref) > mag) mag = fabs(ref);
        }
    }
    printf("%-58s max |D - A.B| = %g (max |A.B| %g)   D[0][0..3] = %g %g %g %g\n", name, worst, mag, h[0], h[1], h[2], h[3]);
    cudaFree(d);
}

int main() {
    //                                            M   N  aMN bMN  aLBO aSBO  bLBO bSBO   aSk  aSmn  bSk  bSmn  aSW bSW es rows32 ak0
    run("tf32 A K none M=128, B K none N=64", Cfg{128, 64, 0, 0, 128, 768, 128, 768, 128, 768, 128, 768, 0, 0, 4, 4, 0});
    run("tf32 A K 128B/32B M=128 k0=0, B K none N=16", Cfg{128, 16, 0, 0, 16, 1024, 128, 768, 0, 1024, 128, 768, 1, 0, 4, 4, 0});
    run("tf32 A K 128B/32B M=128 k0=8, B K none N=16", Cfg{128, 16, 0, 0, 16, 1024, 128, 768, 0, 1024, 128, 768, 1, 0, 4, 4, 8});
    run("tf32 A K 128B/32B M=128 k0=24, B K none N=16", Cfg{128, 16, 0, 0, 16, 1024, 128, 768, 0, 1024, 128, 768, 1, 0, 4, 4, 24});
    run("tf32 A K SW128 M=128 k0=8, B K none N=16", Cfg{128, 16, 0, 0, 16, 1024, 128, 768, 0, 1024, 128, 768, 128, 0, 4, 4, 8});
    run("tf32 A K none, B K none N=16 with LBO 144", Cfg{128, 16, 0, 0, 128, 768, 144, 2304, 128, 768, 144, 2304, 0, 0, 4, 4, 0});
    run("tf32 A MN 128B/32B M=64 (LBO 16K), B K none N=8 LBO 144", Cfg{64, 8, 1, 0, 16384, 512, 144, 4608, 512, 16384, 144, 4608, 1, 0, 4, 4, 0});
    return 0;
}
