// Pose-corrective contraction on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-level accuracy by a 3 x TF32 split.
//
//   C[m][n] = sum_k A[m][k] B[k][n] + Vs[row(m)][n]        M = bodies, N = MH_LD3V (20672), K = MH_KPF (192)
//
// Reference path replaced: `v_posed = v_shaped + matmul(pose_feature, posedirs)` (mhmocap/smpl.py:549-553) -- one of the two
// genuinely dense contractions of the hot path (hard part 2).
//
// One CTA per 128 x 256 output tile.  K is walked in chunks of 32 through two shared-memory stages.  Every value is split into
// x_hi = tf32(x), x_lo = tf32(x - x_hi); both halves sit in the canonical no-swizzle K-major UMMA layout (8 x 16-byte core
// matrices).  The BASIS operand is constant: it is split and laid out per (tile, chunk) once at mh_set_model (`pextF` / `pextB`),
// so a chunk of it is ONE contiguous block that a single TMA bulk copy (`cp.async.bulk`, mbarrier complete_tx) brings in; the
// per-cycle operand (pose features / vertex gradients) is loaded and split by the threads.  ONE thread then issues, per K = 8 step,
// the three products hi*hi + hi*lo + lo*hi as `tcgen05.mma.cta_group::1.kind::tf32` into the same 128-lane fp32 accumulator in
// TMEM and commits the stage to an mbarrier, so the loads of the next chunk overlap the MMAs.  The epilogue reads the accumulator
// with `tcgen05.ld` (one TMEM lane = one body row per thread), adds v_shaped and writes v_posed.
// The dropped lo*lo term is 2^-22 relative: the result differs from the fp32 FMA chain by < 1e-8 m.
#include "mh_ctx.h"

#define TC_BM 128
#define TC_BN 256
#define TC_BK 32
#define TC_THREADS 256
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;                        // 16 KB: one half (hi or lo) of an A chunk
constexpr int TC_B_BYTES = TC_BK * TC_BN * 4;                        // 32 KB: one half of a B chunk
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;      // A_hi A_lo B_hi B_lo
constexpr int TC_SMEM_BYTES = 2 * TC_STAGE_BYTES;
constexpr int TC_TMEM_COLS = 256;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (all >> 4), descriptor version 1
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

// one TMA bulk copy global -> shared, completion (bytes) on an mbarrier that expects them
__device__ __forceinline__ void tc_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}

__device__ __forceinline__ void tc_split4(const float4 v, float4& hi, float4& lo) {
    tc_split(v.x, hi.x, lo.x); tc_split(v.y, hi.y, lo.y); tc_split(v.z, hi.z, lo.z); tc_split(v.w, hi.w, lo.w);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_fwd_tc(const float* __restrict__ Am, const float* __restrict__ Bsplit,
                                                               const float* __restrict__ Vs, float* __restrict__ C, int M, int Npers,
                                                               int per_body_shape) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) unsigned long long bars[5];                // stage 0 / 1 free, accumulator complete, stage 0 / 1 basis landed
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
    const int ncols = min(TC_BN, MH_LD3V - n0);                        // 256, 192 on the last column tile
    const uint32_t sbase = tc_smem_u32(tc_smem);
    const uint32_t bar0 = tc_smem_u32(&bars[0]);
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    uint32_t phase[2] = {0u, 0u};
    constexpr int NIT = MH_KPF / TC_BK;
    float4 av[4];
    auto load_a = [&](int chunk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int id = tid + TC_THREADS * i;
            const int row = id & (TC_BM - 1), k4 = id >> 7;
            av[i] = (m0 + row < M) ? *reinterpret_cast<const float4*>(Am + (size_t)(m0 + row) * MH_KPF + chunk * TC_BK + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_a(0);
    for (int it = 0; it < NIT; ++it) {
        const int s = it & 1;
        unsigned char* st = tc_smem + s * TC_STAGE_BYTES;
        if (it >= 2) { tc_wait(bar0 + 8 * s, phase[s]); phase[s] ^= 1u; }          // the MMAs that read this stage are done
        // B chunk (hi | lo, 64 KB, already split and in the UMMA layout: 16-byte unit (n, k4) at k4 * 256 + n): one TMA bulk copy
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_bulk_load(sbase + s * TC_STAGE_BYTES + 2 * TC_A_BYTES, Bsplit + ((size_t)blockIdx.x * NIT + it) * (2 * TC_B_BYTES / 4), 2 * TC_B_BYTES,
                         bar0 + 24 + 8 * s);
        }
        // A chunk: 128 rows x 32 k, K-major: 16-byte unit (row, k4) at k4 * 128 + row; the values were loaded one iteration ahead
        float4* Ah = reinterpret_cast<float4*>(st);
        float4* Al = reinterpret_cast<float4*>(st + TC_A_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 hi, lo;
            tc_split4(av[i], hi, lo);
            Ah[tid + TC_THREADS * i] = hi; Al[tid + TC_THREADS * i] = lo;
        }
        if (it + 1 < NIT) load_a(it + 1);                                     // in flight across the barrier and the MMA issue
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            tc_wait(bar0 + 24 + 8 * s, (uint32_t)(it >> 1) & 1u);               // the basis chunk has landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_h = sbase + s * TC_STAGE_BYTES, a_l = a_h + TC_A_BYTES;
            const uint32_t b_h = a_h + 2 * TC_A_BYTES, b_l = b_h + TC_B_BYTES;
#pragma unroll
            for (int j = 0; j < TC_BK / 8; ++j) {                               // K = 8 per instruction
                const uint64_t dah = tc_desc(a_h + j * 2 * (TC_BM * 16), TC_BM * 16, 128), dal = tc_desc(a_l + j * 2 * (TC_BM * 16), TC_BM * 16, 128);
                const uint64_t dbh = tc_desc(b_h + j * 2 * (TC_BN * 16), TC_BN * 16, 128), dbl = tc_desc(b_l + j * 2 * (TC_BN * 16), TC_BN * 16, 128);
                tc_mma_tf32(tmem, dah, dbh, idesc, (it | j) != 0);
                tc_mma_tf32(tmem, dah, dbl, idesc, 1u);
                tc_mma_tf32(tmem, dal, dbh, idesc, 1u);
            }
            tc_commit(bar0 + 8 * s);
            if (it == NIT - 1) tc_commit(bar0 + 16);
        }
    }
    tc_wait(bar0 + 16, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 (one body row per thread), columns 128 (w / 4) .. + 127
    {
        const int q = warp & 3, half = warp >> 2;
        const int m = m0 + 32 * q + lane;
        const int srow = per_body_shape ? m : (m % Npers);
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            const int col = 128 * half + 32 * cb;
            if (col >= ncols) break;                                            // warp-uniform
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)col;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < M) {
                const float4* vs = reinterpret_cast<const float4*>(Vs + (size_t)srow * MH_LD3V + n0 + col);
                float4* out = reinterpret_cast<float4*>(C + (size_t)m * MH_LD3V + n0 + col);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = vs[i];
                    out[i] = make_float4(__uint_as_float(r[4 * i]) + v.x, __uint_as_float(r[4 * i + 1]) + v.y,
                                         __uint_as_float(r[4 * i + 2]) + v.z, __uint_as_float(r[4 * i + 3]) + v.w);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

int mh_gemm_fwd_tc(mh_ctx* c, const float* pf, const float* vshaped, float* vposed, int nbodies, int Npers, int per_body_shape,
                   cudaStream_t st) {
    MH_CUDA(c, cudaFuncSetAttribute(k_gemm_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    k_gemm_fwd_tc<<<dim3(mh_cdiv(MH_LD3V, TC_BN), mh_cdiv(nbodies, TC_BM)), TC_THREADS, TC_SMEM_BYTES, st>>>(pf, c->pextF, vshaped, vposed,
                                                                                                         nbodies, Npers, per_body_shape);
    MH_LAUNCHED(c);
    return MH_OK;
}

// -------------------------------------------------------------------------------------------------------------------------
// Backward contraction: Dpart[ks][m][n] = sum_{k in split ks} E[m][k] Bext[n][k]      (dL/dpose_feature and the shape-blend part
// of dL/dbeta, smpl.py:549-553 / :530 backward).  M = bodies, N = MH_NEXT (208 rows of the extended basis), K = MH_LD3V split in
// MH_KSPLIT ranges of 1088.  Both operands are K-major in global memory already.  Same pipeline as the forward kernel.
constexpr int TCB_B_BYTES = MH_NEXT * TC_BK * 4;                      // 26 KB: one half of a B chunk (208 rows x 32 k)
constexpr int TCB_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TCB_B_BYTES;
constexpr int TCB_SMEM_BYTES = 2 * TCB_STAGE_BYTES;
constexpr int TCB_KLEN = MH_LD3V / MH_KSPLIT;
static_assert(MH_LD3V % MH_KSPLIT == 0 && TCB_KLEN % TC_BK == 0 && MH_NEXT % 16 == 0, "split-K must tile the padded row");

__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_bwd_tc(const float* __restrict__ E, const float* __restrict__ Bsplit,
                                                               float* __restrict__ Dpart, int M, int first_body, int nb_total) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) unsigned long long bars[5];                // stage 0 / 1 free, accumulator complete, stage 0 / 1 basis landed
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TC_BM, ks = blockIdx.y;
    const int kbeg = ks * TCB_KLEN;
    const uint32_t sbase = tc_smem_u32(tc_smem);
    const uint32_t bar0 = tc_smem_u32(&bars[0]);
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(MH_NEXT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    uint32_t phase[2] = {0u, 0u};
    constexpr int NIT = TCB_KLEN / TC_BK;
    float4 av[4];
    auto load_a = [&](int kk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int id = tid + TC_THREADS * i;
            const int row = id & (TC_BM - 1), k4 = id >> 7;
            av[i] = (m0 + row < M) ? *reinterpret_cast<const float4*>(E + (size_t)(first_body + m0 + row) * MH_LD3V + kk + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_a(kbeg);
    for (int it = 0; it < NIT; ++it) {
        const int s = it & 1;
        unsigned char* st = tc_smem + s * TCB_STAGE_BYTES;
        const int k0 = kbeg + it * TC_BK;
        if (it >= 2) { tc_wait(bar0 + 8 * s, phase[s]); phase[s] ^= 1u; }
        // B chunk (hi | lo of 208 basis rows x 32 k, already split, 16-byte unit (n, k4) at k4 * 208 + n): one TMA bulk copy
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_bulk_load(sbase + s * TCB_STAGE_BYTES + 2 * TC_A_BYTES, Bsplit + ((size_t)ks * NIT + it) * (2 * TCB_B_BYTES / 4), 2 * TCB_B_BYTES,
                         bar0 + 24 + 8 * s);
        }
        float4* Ah = reinterpret_cast<float4*>(st);
        float4* Al = reinterpret_cast<float4*>(st + TC_A_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 hi, lo;
            tc_split4(av[i], hi, lo);
            Ah[tid + TC_THREADS * i] = hi; Al[tid + TC_THREADS * i] = lo;
        }
        if (it + 1 < NIT) load_a(k0 + TC_BK);                                  // in flight across the barrier and the MMA issue
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            tc_wait(bar0 + 24 + 8 * s, (uint32_t)(it >> 1) & 1u);               // the basis chunk has landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_h = sbase + s * TCB_STAGE_BYTES, a_l = a_h + TC_A_BYTES;
            const uint32_t b_h = a_h + 2 * TC_A_BYTES, b_l = b_h + TCB_B_BYTES;
#pragma unroll
            for (int j = 0; j < TC_BK / 8; ++j) {
                const uint64_t dah = tc_desc(a_h + j * 2 * (TC_BM * 16), TC_BM * 16, 128), dal = tc_desc(a_l + j * 2 * (TC_BM * 16), TC_BM * 16, 128);
                const uint64_t dbh = tc_desc(b_h + j * 2 * (MH_NEXT * 16), MH_NEXT * 16, 128), dbl = tc_desc(b_l + j * 2 * (MH_NEXT * 16), MH_NEXT * 16, 128);
                tc_mma_tf32(tmem, dah, dbh, idesc, (it | j) != 0);
                tc_mma_tf32(tmem, dah, dbl, idesc, 1u);
                tc_mma_tf32(tmem, dal, dbh, idesc, 1u);
            }
            tc_commit(bar0 + 8 * s);
            if (it == NIT - 1) tc_commit(bar0 + 16);
        }
    }
    tc_wait(bar0 + 16, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        // epilogue: 13 blocks of 16 columns; warps 0-3 take blocks 0-6, warps 4-7 blocks 7-12 of their 32 TMEM lanes
        const int q = warp & 3, half = warp >> 2;
        const int m = m0 + 32 * q + lane;
        const int cb0 = half ? 7 : 0, cb1 = half ? MH_NEXT / 16 : 7;
#pragma unroll 1
        for (int cb = cb0; cb < cb1; ++cb) {
            uint32_t r[16];
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * cb);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < M) {
                float4* out = reinterpret_cast<float4*>(Dpart + ((size_t)ks * nb_total + first_body + m) * MH_NEXT + 16 * cb);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    out[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

int mh_gemm_bwd_tc(mh_ctx* c, const float* E, float* dpf_part, int M, int first_body, int nb_total, cudaStream_t st) {
    MH_CUDA(c, cudaFuncSetAttribute(k_gemm_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM_BYTES));
    k_gemm_bwd_tc<<<dim3(mh_cdiv(M, TC_BM), MH_KSPLIT), TC_THREADS, TCB_SMEM_BYTES, st>>>(E, c->pextB, dpf_part, M, first_body, nb_total);
    MH_LAUNCHED(c);
    return MH_OK;
}

// The constant operand of both contractions, split into TF32 hi | lo halves and laid out per (tile, chunk) exactly as the kernels
// stage it in shared memory, so that one chunk is one contiguous block (one TMA bulk copy).
__device__ __forceinline__ float tc_tf32_rna_bits(float x) {       // cvt.rna.tf32.f32: nearest, ties away from zero, 10 mantissa bits kept
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// forward operand: block (column tile nt, chunk it) = [hi: 8 k4 x 256 n units of 4 k | lo: same]; columns beyond MH_LD3V stay 0
__global__ void k_tc_layout_fwd(const float* __restrict__ pext, float* __restrict__ F, int ntile, int nit) {
    const int64_t total = (int64_t)ntile * nit * TC_BK * TC_BN;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(idx % TC_BN);
        int64_t r = idx / TC_BN;
        const int kk = (int)(r % TC_BK);                                 // 4 * k4 + j
        r /= TC_BK;
        const int it = (int)(r % nit), nt = (int)(r / nit);
        const int col = nt * TC_BN + n;
        if (col >= MH_LD3V) continue;
        const float x = pext[(size_t)(it * TC_BK + kk) * MH_LD3V + col];
        const float hi = tc_tf32_rna_bits(x), lo = tc_tf32_rna_bits(x - hi);
        float* blk = F + ((size_t)nt * nit + it) * (2 * TC_B_BYTES / 4);
        const size_t o = (size_t)((kk >> 2) * TC_BN + n) * 4 + (kk & 3);
        blk[o] = hi;
        blk[TC_B_BYTES / 4 + o] = lo;
    }
}

// backward operand: block (split ks, chunk it) = [hi: 8 k4 x 208 basis rows units of 4 k | lo: same]
__global__ void k_tc_layout_bwd(const float* __restrict__ pext, float* __restrict__ B, int nitb) {
    const int64_t total = (int64_t)MH_KSPLIT * nitb * MH_NEXT * TC_BK;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % TC_BK);
        int64_t r = idx / TC_BK;
        const int n = (int)(r % MH_NEXT);
        r /= MH_NEXT;
        const int it = (int)(r % nitb), ks = (int)(r / nitb);
        const float x = pext[(size_t)n * MH_LD3V + ks * TCB_KLEN + it * TC_BK + kk];
        const float hi = tc_tf32_rna_bits(x), lo = tc_tf32_rna_bits(x - hi);
        float* blk = B + ((size_t)ks * nitb + it) * (2 * TCB_B_BYTES / 4);
        const size_t o = (size_t)((kk >> 2) * MH_NEXT + n) * 4 + (kk & 3);
        blk[o] = hi;
        blk[TCB_B_BYTES / 4 + o] = lo;
    }
}

// Both layouts are derived on the DEVICE from the extended basis c->pext that mh_set_model has just uploaded (on the host the two
// strided passes over 4.3 M elements were most of the 50-80 ms of mh_set_model)
int mh_gemm_tc_prepare(mh_ctx* c) {
    const int ntile = mh_cdiv(MH_LD3V, TC_BN), nit = MH_KPF / TC_BK, nitb = TCB_KLEN / TC_BK;
    MH_TRY(mh_alloc_floats(c, &c->pextF, (int64_t)ntile * nit * (2 * TC_B_BYTES / 4)));
    MH_TRY(mh_alloc_floats(c, &c->pextB, (int64_t)MH_KSPLIT * nitb * (2 * TCB_B_BYTES / 4)));
    k_tc_layout_fwd<<<2048, 256>>>(c->pext, c->pextF, ntile, nit);
    k_tc_layout_bwd<<<2048, 256>>>(c->pext, c->pextB, nitb);
    MH_CUDA(c, cudaGetLastError());
    MH_CUDA(c, cudaStreamSynchronize(0));                                  // mh_set_model is a blocking call, as its uploads are
    return MH_OK;
}

// Testing aid: the pose-corrective contraction alone, C = A (M x 192, host) . pext[0:192] (no shape term), with the tensor-core
// kernel (use_tc = 1) or the FP32 SIMT kernel (0); C_host is (M, MH_LD3V).

extern "C" int mh_debug_gemm_fwd(mh_ctx* c, const float* A_host, float* C_host, int32_t M, int32_t use_tc) {
    if (!c || !A_host || !C_host || M <= 0) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    float *A = nullptr, *Z = nullptr, *C = nullptr;
    MH_CUDA(c, mh_dev_alloc((void**)&A, sizeof(float) * (size_t)M * MH_KPF));
    MH_CUDA(c, mh_dev_alloc((void**)&Z, sizeof(float) * MH_LD3V));
    MH_CUDA(c, mh_dev_alloc((void**)&C, sizeof(float) * (size_t)M * MH_LD3V));
    MH_CUDA(c, cudaMemcpy(A, A_host, sizeof(float) * (size_t)M * MH_KPF, cudaMemcpyHostToDevice));
    MH_CUDA(c, cudaMemset(Z, 0, sizeof(float) * MH_LD3V));
    MH_CUDA(c, cudaMemset(C, 0xff, sizeof(float) * (size_t)M * MH_LD3V));
    int rc = use_tc ? mh_gemm_fwd_tc(c, A, Z, C, M, 1, 0, 0) : mh_gemm_fwd_simt(c, A, Z, C, M, 1, 0, 0);
    if (rc == MH_OK) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(C_host, C, sizeof(float) * (size_t)M * MH_LD3V, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "mh_debug_gemm_fwd: %s", cudaGetErrorString(e)); rc = MH_E_CUDA; }
    }
    mh_dev_free(A); mh_dev_free(Z); mh_dev_free(C);
    return rc;
}

// Testing aid: the backward contraction alone, D (M, 208) = E (M, 20672, host) . pext^T, split-K partials reduced on the host.

extern "C" int mh_debug_gemm_bwd(mh_ctx* c, const float* E_host, float* D_host, int32_t M, int32_t use_tc) {
    if (!c || !E_host || !D_host || M <= 0) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    float *E = nullptr, *P = nullptr;
    const size_t np = (size_t)MH_KSPLIT * M * MH_NEXT;
    MH_CUDA(c, mh_dev_alloc((void**)&E, sizeof(float) * (size_t)M * MH_LD3V));
    MH_CUDA(c, mh_dev_alloc((void**)&P, sizeof(float) * np));
    MH_CUDA(c, cudaMemcpy(E, E_host, sizeof(float) * (size_t)M * MH_LD3V, cudaMemcpyHostToDevice));
    MH_CUDA(c, cudaMemset(P, 0xff, sizeof(float) * np));
    int rc = use_tc ? mh_gemm_bwd_tc(c, E, P, M, 0, M, 0) : mh_gemm_bwd_simt(c, E, P, M, 0, M, 0);
    if (rc == MH_OK) {
        std::vector<float> h(np);
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(h.data(), P, sizeof(float) * np, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "mh_debug_gemm_bwd: %s", cudaGetErrorString(e)); rc = MH_E_CUDA; }
        else
            for (size_t i = 0; i < (size_t)M * MH_NEXT; ++i) {
                float a = 0.f;
                for (int ks = 0; ks < MH_KSPLIT; ++ks) a += h[(size_t)ks * M * MH_NEXT + i];
                D_host[i] = a;
            }
    }
    mh_dev_free(E); mh_dev_free(P);
    return rc;
}
