// Development probe, standalone:  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probe/tc_probe tools/probe/tc_probe.cu
// then  gpurun -- ./tools/probe/tc_probe .  One tcgen05.mma kind::tf32 M=128 N=256 K=8 with
// operands laid out by hand, under several descriptor hypotheses.  Prints which hypothesis reproduces A.B.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Hyp { uint32_t lboA, sboA, lboB, sboB; int bmajor; int layB; };   // layB: 0 = my N-major core matrices, 1 = K-major B (n rows x 8 k)

__global__ void __launch_bounds__(128, 1) k_probe(const float* A, const float* B, float* C, Hyp h, int with_mma, int* flag) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float4* As = reinterpret_cast<float4*>(sm);                 // 128 x 8 tf32 = 4 KB
    float4* Bs = reinterpret_cast<float4*>(sm + 8192);          // 8 x 256 tf32 = 8 KB
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A (r, k): K-major core matrices: 16-byte unit (r, k4) at k4 * 128 + r
    for (int id = tid; id < 256; id += 128) {
        const int r = id & 127, k4 = id >> 7;
        As[id] = make_float4(A[r * 8 + k4 * 4], A[r * 8 + k4 * 4 + 1], A[r * 8 + k4 * 4 + 2], A[r * 8 + k4 * 4 + 3]);
    }
    if (h.layB == 0) {
        // B (k, n) N-major: 16-byte unit (n4, k) at k + n4 * 8
        for (int id = tid; id < 512; id += 128) {
            const int k = id & 7, n4 = id >> 3;
            Bs[id] = make_float4(B[k * 256 + n4 * 4], B[k * 256 + n4 * 4 + 1], B[k * 256 + n4 * 4 + 2], B[k * 256 + n4 * 4 + 3]);
        }
    } else {
        // B as K-major (n rows, k contiguous): 16-byte unit (n, k4) at k4 * 256 + n
        for (int id = tid; id < 512; id += 128) {
            const int n = id & 255, k4 = id >> 8;
            Bs[id] = make_float4(B[(k4 * 4) * 256 + n], B[(k4 * 4 + 1) * 256 + n], B[(k4 * 4 + 2) * 256 + n], B[(k4 * 4 + 3) * 256 + n]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) flag[0] = (int)tmem;
    if (with_mma) {
        if (tid == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)h.bmajor << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t da = (uint64_t)((s32(As) & 0x3FFFFu) >> 4) | ((uint64_t)(h.lboA >> 4) << 16) | ((uint64_t)(h.sboA >> 4) << 32) | (1ull << 46);
            const uint64_t db = (uint64_t)((s32(Bs) & 0x3FFFFu) >> 4) | ((uint64_t)(h.lboB >> 4) << 16) | ((uint64_t)(h.sboB >> 4) << 32) | (1ull << 46);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        }
        uint32_t done = 0; int spins = 0;
        while (!done && spins < (1 << 22)) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
            ++spins;
        }
        if (tid == 0) flag[1] = spins;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    } else {
        // no MMA: write a pattern into TMEM with tcgen05.st, to validate the ld path alone
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(100 * (32 * warp + lane) + i));
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    for (int cb = 0; cb < 8; ++cb) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + 32 * cb;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) C[(32 * warp + lane) * 256 + 32 * cb + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<float> A(128 * 8), B(8 * 256), C(128 * 256), R(128 * 256);
    srand(1);
    for (auto& v : A) v = (float)((rand() % 17) - 8);            // small integers: exact in tf32
    for (auto& v : B) v = (float)((rand() % 13) - 6);
    for (int r = 0; r < 128; ++r) for (int n = 0; n < 256; ++n) { float s = 0; for (int k = 0; k < 8; ++k) s += A[r * 8 + k] * B[k * 256 + n]; R[r * 256 + n] = s; }
    float *dA, *dB, *dC; int* dF;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4); cudaMalloc(&dF, 16);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    // ld/st path alone
    cudaMemset(dC, 0xff, C.size() * 4); cudaMemset(dF, 0, 16);
    k_probe<<<1, 128, 32768>>>(dA, dB, dC, Hyp{0, 0, 0, 0, 0, 0}, 0, dF);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    int F[4]; cudaMemcpy(F, dF, 16, cudaMemcpyDeviceToHost);
    printf("st/ld: %s tmem 0x%x  C[0][0..3] %g %g %g %g  C[1][0] %g C[33][2] %g C[127][7] %g (expect 0 1 2 3 | 100 | 3302 | 12707)\n", cudaGetErrorString(e), F[0],
           C[0], C[1], C[2], C[3], C[256], C[33 * 256 + 2], C[127 * 256 + 7]);
    const Hyp hyps[] = {
        {2048, 128, 8192, 128, 1, 0},    // mine: A lbo = K-chunk stride, sbo = 8-row stride; B N-major lbo = k-group, sbo = n4 stride
        {128, 2048, 128, 8192, 1, 0},    // LBO / SBO swapped
        {2048, 128, 128, 8192, 1, 0},    // B swapped only
        {128, 2048, 8192, 128, 1, 0},    // A swapped only
        {2048, 128, 4096, 128, 0, 1},    // B K-major (n rows): lbo = K-chunk stride (256 * 16), sbo = 8-row stride
    };
    for (size_t i = 0; i < sizeof(hyps) / sizeof(hyps[0]); ++i) {
        cudaMemset(dC, 0xff, C.size() * 4); cudaMemset(dF, 0, 16);
        k_probe<<<1, 128, 32768>>>(dA, dB, dC, hyps[i], 1, dF);
        e = cudaDeviceSynchronize();
        cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(F, dF, 16, cudaMemcpyDeviceToHost);
        int bad = 0, zero = 0; for (size_t j = 0; j < C.size(); ++j) { bad += C[j] != R[j]; zero += C[j] == 0.f; }
        printf("hyp %zu: %s spins %d mismatches %d / %zu zeros %d   C[0][0..3] %g %g %g %g expect %g %g %g %g | C[9][5] %g expect %g\n", i, cudaGetErrorString(e), F[1], bad, C.size(), zero,
               C[0], C[1], C[2], C[3], R[0], R[1], R[2], R[3], C[9 * 256 + 5], R[9 * 256 + 5]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
