// Per-frame image planes: ingest-time compaction of the instance masks, the derived constants of the raster
// terms, the depth-order prepass, and device-side synthesis of test / bench inputs.
//
// Reference code replaced (mhmocap/optimizer.py): the per-cycle H2D copy of every modality (:396-397) becomes a
// one-time ingest; erode(erode(seg_mask)) (:306-309, 434; morphology.py:23-33) is evaluated once because its
// input is constant; mask / pose validity flags (:404-409); the near-to-far person order (:450) and the
// occlusion accumulator (:458, 475) reduced to per-frame pixel counts.
#include "mh_ctx.h"

// seg (count, N, HW) f32 {0,1} (as the reference dataset delivers it, utils.py:329-331) or u8 / bool {0,1} -> one 32-bit plane
// per frame, bit n = person n covers the pixel
template <typename T>
__global__ void k_compact(const T* __restrict__ seg, int N, int64_t HW, uint32_t* __restrict__ cbits, int* __restrict__ flags) {
    const int t = blockIdx.y;
    bool bad = false;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        uint32_t bits = 0;
        for (int n = 0; n < N; ++n) {
            const T v = seg[((int64_t)t * N + n) * HW + p];
            if (v != (T)0) { bits |= 1u << n; bad |= (v != (T)1); }
        }
        cbits[(int64_t)t * HW + p] = bits;
    }
    if (bad) atomicOr(flags, 1);
}

int mh_ingest_compact(mh_ctx* c, int t0, int count, int seg_is_u8, cudaStream_t st) {
    const int64_t HW = (int64_t)c->d.H * c->d.W;
    const dim3 grid(std::min(mh_cdiv(HW, 256), 1024), count);
    if (seg_is_u8) k_compact<uint8_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(c->stage), c->d.N, HW, c->cbits + (int64_t)t0 * HW, c->devflags);
    else k_compact<float><<<grid, 256, 0, st>>>(c->stage, c->d.N, HW, c->cbits + (int64_t)t0 * HW, c->devflags);
    MH_LAUNCHED(c);
    return MH_OK;
}

// ebits = AND over the 5x5 window of cbits; out-of-image neighbours do not erode (morphology.py:29-31 pads
// the "x < 0.5" map with zeros).  Also the per-person mask areas.
__global__ void k_erode_area(const uint32_t* __restrict__ cbits, int H, int W, int N, uint32_t* __restrict__ ebits, int* __restrict__ area) {
    __shared__ int sarea[MH_MAXN];
    const int t = blockIdx.y;
    if (threadIdx.x < MH_MAXN) sarea[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t* cb = cbits + (int64_t)t * H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
        const int y = p / W, x = p - y * W;
        uint32_t e = 0xffffffffu;
        for (int dy = -2; dy <= 2; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -2; dx <= 2; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                e &= cb[yy * W + xx];
            }
        }
        ebits[(int64_t)t * H * W + p] = e;
        uint32_t b = cb[p];
        while (b) { const int n = __ffs(b) - 1; b &= b - 1; atomicAdd(&sarea[n], 1); }
    }
    __syncthreads();
    if (threadIdx.x < N && sarea[threadIdx.x]) atomicAdd(&area[t * N + threadIdx.x], sarea[threadIdx.x]);
}

__global__ void k_validity(const float* __restrict__ pose2d, const int* __restrict__ area, int TN, float thr, float min_area,
                           uint8_t* __restrict__ pose2d_valid, uint8_t* __restrict__ mask_valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TN) return;
    int cnt = 0;
    for (int k = 0; k < MH_NJR; ++k) cnt += pose2d[((size_t)i * MH_NJR + k) * 3 + 2] >= thr;
    pose2d_valid[i] = cnt >= 2;                                 // optimizer.py:404-405
    mask_valid[i] = (float)area[i] >= min_area;                 // optimizer.py:407-409
}

int mh_ingest_derive(mh_ctx* c, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int TN = d.T * d.N;
    MH_CUDA(c, cudaMemsetAsync(c->maskarea, 0, sizeof(int) * TN, st));
    k_erode_area<<<dim3(std::min(mh_cdiv((int64_t)d.H * d.W, 256), 512), d.T), 256, 0, st>>>(c->cbits, d.H, d.W, d.N, c->ebits, c->maskarea);
    MH_LAUNCHED(c);
    k_validity<<<mh_cdiv(TN, 128), 128, 0, st>>>(c->pose2d, c->maskarea, TN, c->c.joint_confidence_thr, (float)(0.005 * d.H * d.W),
                                                 c->pose2d_valid, c->mask_valid);
    MH_LAUNCHED(c);
    // the order-dependent counts must be rebuilt for the new planes
    MH_CUDA(c, cudaMemsetAsync(c->order, 0xff, sizeof(int) * TN, st));
    return MH_OK;
}

__global__ void k_expand(const uint32_t* __restrict__ cbits, int N, int64_t HW, float* __restrict__ seg) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = cbits[p];
        for (int n = 0; n < N; ++n) seg[(int64_t)n * HW + p] = (float)((b >> n) & 1u);
    }
}

int mh_expand_planes(mh_ctx* c, int t, float* seg_dev, cudaStream_t st) {
    const int64_t HW = (int64_t)c->d.H * c->d.W;
    k_expand<<<std::min(mh_cdiv(HW, 256), 2048), 256, 0, st>>>(c->cbits + (int64_t)t * HW, c->d.N, HW, seg_dev);
    MH_LAUNCHED(c);
    return MH_OK;
}

// -------------------------------------------------------------------------------------------------
// near -> far person order per frame from the optimised translations (optimizer.py:450); frames whose order
// changed are marked dirty and their counts reset
__global__ void k_order(const float* __restrict__ poses_T, int T, int N, int* __restrict__ order, uint32_t* __restrict__ premask,
                        int* __restrict__ dirty, int* __restrict__ rankcnt) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int ord[MH_MAXN];
    float z[MH_MAXN];
    for (int n = 0; n < N; ++n) {                               // stable insertion sort, ascending z
        const float zn = poses_T[((size_t)t * N + n) * 3 + 2];
        int k = n;
        while (k > 0 && z[k - 1] > zn) { z[k] = z[k - 1]; ord[k] = ord[k - 1]; --k; }
        z[k] = zn; ord[k] = n;
    }
    bool changed = false;
    uint32_t pre = 0;
    for (int k = 0; k < N; ++k) {
        changed |= order[t * N + k] != ord[k];
        order[t * N + k] = ord[k];
        premask[t * N + k] = pre;
        pre |= 1u << ord[k];
    }
    dirty[t] = changed;
    if (changed) for (int k = 0; k <= N; ++k) rankcnt[t * (N + 1) + k] = 0;
}

// rankcnt[t][q] = pixels whose FIRST covering person (in depth order) sits at position q; [N] = uncovered.
// From these: sum(1 - acc_q) = sum_{q' >= q} rankcnt[q'] + rankcnt[N], and the alpha = 0 silhouette energy of
// position q is rankcnt[q] (the masks are binary).
__global__ void k_rank_count(const uint32_t* __restrict__ cbits, const int* __restrict__ order, const int* __restrict__ dirty, int N,
                             int64_t HW, int* __restrict__ rankcnt) {
    const int t = blockIdx.y;
    if (!dirty[t]) return;
    __shared__ int pos[MH_MAXN];
    __shared__ int hist[MH_MAXN + 1];
    if (threadIdx.x < N) pos[order[t * N + threadIdx.x]] = threadIdx.x;
    if (threadIdx.x <= N) hist[threadIdx.x] = 0;
    __syncthreads();
    int none = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        uint32_t b = cbits[(int64_t)t * HW + p];
        if (!b) { ++none; continue; }
        int best = N;
        while (b) { const int n = __ffs(b) - 1; b &= b - 1; best = min(best, pos[n]); }
        atomicAdd(&hist[best], 1);
    }
    if (none) atomicAdd(&hist[N], none);
    __syncthreads();
    if (threadIdx.x <= N && hist[threadIdx.x]) atomicAdd(&rankcnt[t * (N + 1) + threadIdx.x], hist[threadIdx.x]);
}

int mh_render_prepass(mh_ctx* c, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int64_t HW = (int64_t)d.H * d.W;
    k_order<<<mh_cdiv(d.T, 64), 64, 0, st>>>(c->params + c->off[MH_P_POSES_T], d.T, d.N, c->order, c->premask, c->dirty, c->rankcnt);
    MH_LAUNCHED(c);
    k_rank_count<<<dim3(std::min(mh_cdiv(HW, 1024), 64), d.T), 256, 0, st>>>(c->cbits, c->order, c->dirty, d.N, HW, c->rankcnt);
    MH_LAUNCHED(c);
    return MH_OK;
}

// -------------------------------------------------------------------------------------------------
// Input synthesis on the device (bench / tests; oracle.synth.assemble_inputs semantics): hard z-buffers of the
// CURRENT parameters -> nearest-person-wins instance bits and the min-max normalised disparity of
// (ground plane y = y_ground) U (wall z = z_wall) U persons.
__device__ __forceinline__ unsigned f2ord(float f) { return __float_as_uint(f); }      // positive floats order like their bits

__global__ void k_synth_compose(const float* __restrict__ zb, int N, int H, int W, float fy, float cy, float y_ground, float z_wall,
                                uint32_t* __restrict__ cbits, float* __restrict__ disp, unsigned* __restrict__ mm) {
    unsigned lo = 0x7f800000u, hi = 0u;
    const int64_t HW = (int64_t)H * W;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / W);
        float zmin = INFINITY; int who = -1;
        for (int n = 0; n < N; ++n) { const float z = zb[(int64_t)n * HW + p]; if (z > 0.f && z < zmin) { zmin = z; who = n; } }
        const float v = ((float)y + 0.5f - cy) / fy;
        const float zg = v > 1e-6f ? y_ground / v : INFINITY;
        const float scene = fminf(zg, z_wall);
        const float depth = who >= 0 ? fminf(zmin, scene) : scene;
        const float dsp = 1.0f / depth;
        cbits[p] = who >= 0 ? (1u << who) : 0u;
        disp[p] = dsp;
        lo = min(lo, f2ord(dsp)); hi = max(hi, f2ord(dsp));
    }
    for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}

__global__ void k_synth_normalize(float* __restrict__ disp, int64_t HW, const unsigned* __restrict__ mm) {
    const float lo = __uint_as_float(mm[0]), hi = __uint_as_float(mm[1]);
    const float inv = 1.0f / fmaxf(hi - lo, 1e-9f);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) disp[p] = (disp[p] - lo) * inv;
}

int mh_render_planes(mh_ctx* c, int t, int n, float* zbuf_dev, float* alpha_dev, float blur_d, float blur_s, cudaStream_t st);

int mh_render_synth(mh_ctx* c, float y_ground, float z_wall, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int64_t HW = (int64_t)d.H * d.W;
    float* zb; unsigned* mm;
    MH_CUDA(c, mh_dev_alloc((void**)&zb, sizeof(float) * (d.N + 1) * HW));
    MH_CUDA(c, mh_dev_alloc((void**)&mm, sizeof(unsigned) * 2 * d.T));
    int r = MH_OK;
    std::vector<unsigned> init(2 * d.T);
    for (int t = 0; t < d.T; ++t) { init[2 * t] = 0x7f800000u; init[2 * t + 1] = 0u; }
    if (cudaMemcpyAsync(mm, init.data(), sizeof(unsigned) * 2 * d.T, cudaMemcpyHostToDevice, st) != cudaSuccess) r = MH_E_CUDA;
    for (int t = 0; t < d.T && r == MH_OK; ++t) {
        for (int n = 0; n < d.N && r == MH_OK; ++n) r = mh_render_planes(c, t, n, zb + (int64_t)n * HW, zb + (int64_t)d.N * HW, 0.f, 0.f, st);
        if (r != MH_OK) break;
        k_synth_compose<<<std::min(mh_cdiv(HW, 256), 1024), 256, 0, st>>>(zb, d.N, d.H, d.W, c->K[4], c->K[5], y_ground, z_wall,
                                                                           c->cbits + (int64_t)t * HW, c->depth + (int64_t)t * HW, mm + 2 * t);
        c->launches++;
        k_synth_normalize<<<std::min(mh_cdiv(HW, 256), 1024), 256, 0, st>>>(c->depth + (int64_t)t * HW, HW, mm + 2 * t);
        c->launches++;
    }
    cudaStreamSynchronize(st);
    mh_dev_free(zb); mh_dev_free(mm);
    if (r != MH_OK) return r;
    MH_CUDA(c, cudaGetLastError());
    return MH_OK;
}
