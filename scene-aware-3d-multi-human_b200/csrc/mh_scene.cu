// Scene-geometry aggregation on the device: masked temporal median of the per-frame scene depths (and, once at the end
// of fit(), of the RGB frames) -- the host work that dominates the reference's fit() from cycle 30 on.
//
// Reference code replaced: aggegrate_scene_geometry_median (mhmocap/fhsog.py:180-202) called every cycle >= 30
// (mhmocap/optimizer.py:578-582) on depths = 1 / target_disp (optimizer.py:425-426), images and backmasks collected from
// the dataloader (optimizer.py:399-400).  np.ma.median semantics: per pixel, over the frames whose backmask is non-zero,
// the middle value (odd count) or the float32 mean of the two middle values (even count); a pixel that is never
// background gets data 0 and mask False.
//
// Exact selection without sorting: a radix-16 descent over the float bit patterns (positive floats order like their
// bits).  Each pass streams the T frames once (coalesced over pixels) and histograms one 4-bit digit of the values that
// match the prefix found so far; the per-pixel histograms are plain sums over frames, so with frame-sharded ranks the
// caller all-reduces them (SUM) between passes and every rank descends identically -- no transpose of the (T,H,W)
// volume is needed.  10 passes for the depth median (count, 8 digits, upper neighbour), 4 for the image median.
#include "mh_ctx.h"

struct MhSceneState {
    uint8_t* back;          // (T, HW)   backmask != 0
    uint8_t* images;        // (T, HW, 3) or null
    float* hist;            // (16, HW)  per-pass histogram (floats: counts <= T are exact, and the caller's all-reduce is float)
    float* aux;             // (6, HW)   last pass: planes 0-2 count <= lower median (per channel), planes 3-5 smallest value above it
    uint32_t* prefix;       // (3, HW)   bits found so far (depth: plane 0; image: one plane per channel)
    uint32_t* krem;         // (3, HW)   remaining rank inside the prefix class
    uint32_t* ntot;         // (HW)      valid frames per pixel
    float* ab;              // (T, 2)    per-frame a = 1/min_z - 1/max_z, b = 1/max_z
    float* out_depth; uint8_t* out_mask; uint8_t* out_img;      // (HW), (HW), (HW,3)
    int64_t HW;
    bool has_back, has_images;
};

static MhSceneState* scene_state(mh_ctx* c) { return reinterpret_cast<MhSceneState*>(c->scene_state); }

extern "C" int mh_scene_set_back(mh_ctx* c, int32_t t0, int32_t count, const uint8_t* back_host, const uint8_t* images_host, void* stream) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    const mh_dims& d = c->d;
    if (t0 < 0 || count < 1 || t0 + count > d.T || !back_host) MH_FAIL(c, MH_E_ARG, "mh_scene_set_back: bad arguments");
    const int64_t HW = (int64_t)d.H * d.W;
    if (!c->scene_state) {
        MhSceneState* s = new MhSceneState();
        memset(s, 0, sizeof(*s));
        s->HW = HW;
        c->scene_state = s;
        cudaError_t e = mh_dev_alloc((void**)&s->back, (size_t)d.T * HW);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->hist, sizeof(float) * 16 * HW * 3);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->aux, sizeof(float) * 2 * HW * 3);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->prefix, sizeof(uint32_t) * 3 * HW);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->krem, sizeof(uint32_t) * 3 * HW);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->ntot, sizeof(uint32_t) * HW);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->ab, sizeof(float) * 2 * d.T);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->out_depth, sizeof(float) * HW);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->out_mask, HW);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->out_img, 3 * HW);
        if (e == cudaSuccess) e = cudaMemset(s->back, 0, (size_t)d.T * HW);
        if (e != cudaSuccess) MH_FAIL(c, MH_E_CUDA, "scene state: %s", cudaGetErrorString(e));
    }
    MhSceneState* s = scene_state(c);
    cudaStream_t st = (cudaStream_t)stream;
    MH_CUDA(c, cudaMemcpyAsync(s->back + (size_t)t0 * HW, back_host, (size_t)count * HW, cudaMemcpyHostToDevice, st));
    s->has_back = true;
    if (images_host) {
        if (!s->images) MH_CUDA(c, mh_dev_alloc((void**)&s->images, (size_t)d.T * HW * 3));
        MH_CUDA(c, cudaMemcpyAsync(s->images + (size_t)t0 * HW * 3, images_host, (size_t)count * HW * 3, cudaMemcpyHostToDevice, st));
        s->has_images = true;
    }
    return MH_OK;
}

void mh_scene_free(mh_ctx* c) {
    MhSceneState* s = scene_state(c);
    if (!s) return;
    mh_dev_free(s->back); if (s->images) mh_dev_free(s->images);
    mh_dev_free(s->hist); mh_dev_free(s->aux); mh_dev_free(s->prefix); mh_dev_free(s->krem); mh_dev_free(s->ntot); mh_dev_free(s->ab);
    mh_dev_free(s->out_depth); mh_dev_free(s->out_mask); mh_dev_free(s->out_img);
    delete s;
    c->scene_state = nullptr;
}

// per-frame scale / offset of target_disp = d * (1/min_z - 1/max_z) + 1/max_z  (optimizer.py:425, 683-688)
__global__ void k_scene_ab(const float* __restrict__ zmin_lin, const float* __restrict__ zmax_lin, int T, float* __restrict__ ab) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float minz = logf(1.0f + expf(zmin_lin[t]));
    const float maxz = minz + 1.0f + logf(1.0f + expf(zmax_lin[t]));
    const float ia = __fdiv_rn(1.0f, minz), ib = __fdiv_rn(1.0f, maxz);
    ab[2 * t] = __fsub_rn(ia, ib);
    ab[2 * t + 1] = ib;
}

__device__ __forceinline__ uint32_t depth_bits(float d, float a, float b) {
    return __float_as_uint(__fdiv_rn(1.0f, __fadd_rn(__fmul_rn(d, a), b)));        // 1 / target_disp, rounded like torch
}

// The passes over the T frames of a pixel are bound by memory LATENCY when every iteration waits for its own mask byte and then for
// its depth (one dependent pair in flight per thread: measured 4x off the HBM bound).  MED_U frames are therefore loaded first, both
// planes unconditionally, with volatile loads the compiler may neither sink below the mask test nor reorder, and judged afterwards.
#define MED_U 8
__device__ __forceinline__ float ld_f32_issue(const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ unsigned ld_u8_issue(const uint8_t* p) { unsigned v; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

// ---- depth passes ---------------------------------------------------------------------------------------------------
// pass 0: valid-frame count.  hist[0] = local count.
__global__ void k_med_count(const uint8_t* __restrict__ back, int T, int64_t HW, float* __restrict__ hist) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    int n = 0;
#pragma unroll 8
    for (int t = 0; t < T; ++t) n += back[(size_t)t * HW + p] != 0;
    hist[p] = (float)n;
}

// after the all-reduce of pass 0: rank of the LOWER median inside the whole set
__global__ void k_med_start(const float* __restrict__ hist, int64_t HW, uint32_t* __restrict__ ntot, uint32_t* __restrict__ prefix,
                            uint32_t* __restrict__ krem, int planes) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const uint32_t n = (uint32_t)hist[p];
    ntot[p] = n;
    for (int c = 0; c < planes; ++c) { prefix[(size_t)c * HW + p] = 0u; krem[(size_t)c * HW + p] = n ? (n - 1) / 2 : 0u; }
}

// digit pass: histogram of digit `shift` over the values whose higher bits equal the prefix
__global__ void k_med_digit_depth(const float* __restrict__ depth, const uint8_t* __restrict__ back, const float* __restrict__ ab, int T,
                                  int64_t HW, const uint32_t* __restrict__ prefix, int shift, float* __restrict__ hist) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const uint32_t pre = prefix[p];
    const uint32_t himask = shift >= 28 ? 0u : (0xffffffffu << (shift + 4));
    unsigned long long lo = 0ull, hi = 0ull;            // 16 counters of 8 bits would overflow: two words of 4 x 16-bit lanes per half
    unsigned long long lo2 = 0ull, hi2 = 0ull;
    for (int t0 = 0; t0 < T; t0 += MED_U) {
        unsigned bk[MED_U];
        float dv[MED_U];
#pragma unroll
        for (int j = 0; j < MED_U; ++j) {
            const size_t o = (size_t)min(t0 + j, T - 1) * HW + p;
            bk[j] = ld_u8_issue(back + o);
            dv[j] = ld_f32_issue(depth + o);
        }
#pragma unroll
        for (int j = 0; j < MED_U; ++j) {
            const int t = t0 + j;
            if (t >= T || !bk[j]) continue;
            const uint32_t b = depth_bits(dv[j], ab[2 * t], ab[2 * t + 1]);
            if ((b & himask) != pre) continue;
            const uint32_t dg = (b >> shift) & 15u;
            const unsigned long long one = 1ull << ((dg & 3u) * 16);
            switch (dg >> 2) { case 0: lo += one; break; case 1: hi += one; break; case 2: lo2 += one; break; default: hi2 += one; }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        hist[(size_t)k * HW + p] = (float)((lo >> (16 * k)) & 0xffffull);
        hist[(size_t)(4 + k) * HW + p] = (float)((hi >> (16 * k)) & 0xffffull);
        hist[(size_t)(8 + k) * HW + p] = (float)((lo2 >> (16 * k)) & 0xffffull);
        hist[(size_t)(12 + k) * HW + p] = (float)((hi2 >> (16 * k)) & 0xffffull);
    }
}

// after the all-reduce of a digit pass: descend into the digit that holds the wanted rank
__global__ void k_med_descend(const float* __restrict__ hist, int64_t HW, int shift, uint32_t* __restrict__ prefix, uint32_t* __restrict__ krem) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    uint32_t k = krem[p];
    uint32_t dg = 15u;
    for (uint32_t d = 0; d < 16u; ++d) {
        const uint32_t c = (uint32_t)hist[(size_t)d * HW + p];
        if (k < c) { dg = d; break; }
        k -= c;
    }
    prefix[p] |= dg << shift;
    krem[p] = k;
}

// upper-neighbour pass: aux[0] = count of values <= lower median, aux[1] = smallest value above it
__global__ void k_med_upper_depth(const float* __restrict__ depth, const uint8_t* __restrict__ back, const float* __restrict__ ab, int T,
                                  int64_t HW, const uint32_t* __restrict__ prefix, float* __restrict__ aux) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const uint32_t m = prefix[p];
    int le = 0;
    uint32_t above = 0x7f800000u;
    for (int t0 = 0; t0 < T; t0 += MED_U) {
        unsigned bk[MED_U];
        float dv[MED_U];
#pragma unroll
        for (int j = 0; j < MED_U; ++j) {
            const size_t o = (size_t)min(t0 + j, T - 1) * HW + p;
            bk[j] = ld_u8_issue(back + o);
            dv[j] = ld_f32_issue(depth + o);
        }
#pragma unroll
        for (int j = 0; j < MED_U; ++j) {
            const int t = t0 + j;
            if (t >= T || !bk[j]) continue;
            const uint32_t b = depth_bits(dv[j], ab[2 * t], ab[2 * t + 1]);
            if (b <= m) ++le; else above = min(above, b);
        }
    }
    aux[p] = (float)le;
    aux[3 * HW + p] = __uint_as_float(above);
}

__global__ void k_med_finish_depth(const float* __restrict__ aux, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ ntot,
                                   int64_t HW, float* __restrict__ out_depth, uint8_t* __restrict__ out_mask) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const uint32_t n = ntot[p];
    if (n == 0) { out_depth[p] = 0.f; out_mask[p] = 0; return; }
    const float lo = __uint_as_float(prefix[p]);
    float hi = lo;
    if ((n & 1u) == 0u) {                                          // even count: the upper median is the next order statistic
        const uint32_t le = (uint32_t)aux[p];
        if (le < n / 2 + 1) hi = aux[3 * HW + p];
    }
    out_depth[p] = __fdiv_rn(__fadd_rn(lo, hi), 2.0f);
    out_mask[p] = 1;
}

// ---- image passes (u8, three channels at once; 2 digit passes + upper neighbour) ---------------------------------------
__global__ void k_med_digit_img(const uint8_t* __restrict__ img, const uint8_t* __restrict__ back, int T, int64_t HW,
                                const uint32_t* __restrict__ prefix, int shift, float* __restrict__ hist) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    uint32_t pre[3] = {prefix[p], prefix[HW + p], prefix[2 * HW + p]};
    unsigned short cnt[3][16];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 16; ++k) cnt[c][k] = 0;
    const uint32_t himask = shift == 4 ? 0u : 0xf0u;
    for (int t = 0; t < T; ++t) {
        if (!back[(size_t)t * HW + p]) continue;
        const uint8_t* px = img + ((size_t)t * HW + p) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t b = px[c];
            if ((b & himask) != pre[c]) continue;
            const uint32_t dg = (b >> shift) & 15u;
#pragma unroll
            for (int k = 0; k < 16; ++k) cnt[c][k] += (dg == (uint32_t)k);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 16; ++k) hist[((size_t)c * 16 + k) * HW + p] = (float)cnt[c][k];
}

__global__ void k_med_descend_img(const float* __restrict__ hist, int64_t HW, int shift, uint32_t* __restrict__ prefix, uint32_t* __restrict__ krem) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    for (int c = 0; c < 3; ++c) {
        uint32_t k = krem[(size_t)c * HW + p];
        uint32_t dg = 15u;
        for (uint32_t d = 0; d < 16u; ++d) {
            const uint32_t n = (uint32_t)hist[((size_t)c * 16 + d) * HW + p];
            if (k < n) { dg = d; break; }
            k -= n;
        }
        prefix[(size_t)c * HW + p] |= dg << shift;
        krem[(size_t)c * HW + p] = k;
    }
}

__global__ void k_med_upper_img(const uint8_t* __restrict__ img, const uint8_t* __restrict__ back, int T, int64_t HW,
                                const uint32_t* __restrict__ prefix, float* __restrict__ aux) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const uint32_t m[3] = {prefix[p], prefix[HW + p], prefix[2 * HW + p]};
    int le[3] = {0, 0, 0};
    uint32_t above[3] = {1024u, 1024u, 1024u};
    for (int t = 0; t < T; ++t) {
        if (!back[(size_t)t * HW + p]) continue;
        const uint8_t* px = img + ((size_t)t * HW + p) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) { const uint32_t b = px[c]; if (b <= m[c]) ++le[c]; else above[c] = min(above[c], b); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { aux[(size_t)c * HW + p] = (float)le[c]; aux[(size_t)(3 + c) * HW + p] = (float)above[c]; }
}

__global__ void k_med_finish_img(const float* __restrict__ aux, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ ntot, int64_t HW,
                                 uint8_t* __restrict__ out_img) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const uint32_t n = ntot[p];
    for (int c = 0; c < 3; ++c) {
        uint32_t lo = prefix[(size_t)c * HW + p], hi = lo;
        if (n == 0) { out_img[p * 3 + c] = 0; continue; }
        if ((n & 1u) == 0u && (uint32_t)aux[(size_t)c * HW + p] < n / 2 + 1) hi = (uint32_t)aux[(size_t)(3 + c) * HW + p];
        out_img[p * 3 + c] = (uint8_t)((lo + hi) >> 1);             // float64 mean truncated by astype(uint8) (fhsog.py:191)
    }
}

// ---- pass driver ----------------------------------------------------------------------------------------------------
// which: 0 depth, 1 image.  The caller runs, for pass = 0 .. n_passes-1:  mh_scene_median_pass ; all-reduce(SUM) of the
// MH_BUF_MEDIAN_HIST view (pass 0 .. n-2) resp. all-reduce of MH_BUF_MEDIAN_AUX (last pass: plane 0 SUM, plane 1 MIN) ;
// then mh_scene_median_finish.  Depth: 10 passes (count, 8 digits, upper).  Image: 4 passes (count, 2 digits, upper).
extern "C" int mh_scene_median_pass(mh_ctx* c, int32_t which, int32_t pass, void* stream) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    MhSceneState* s = scene_state(c);
    if (!s || !s->has_back) MH_FAIL(c, MH_E_STATE, "mh_scene_median_pass: mh_scene_set_back first");
    if (which == 1 && !s->has_images) MH_FAIL(c, MH_E_STATE, "mh_scene_median_pass: no images were given");
    cudaStream_t st = (cudaStream_t)stream;
    const mh_dims& d = c->d;
    const int64_t HW = s->HW;
    const int grid = mh_cdiv(HW, 256);
    const int npass = which == 0 ? 10 : 4;
    if (pass < 0 || pass >= npass) MH_FAIL(c, MH_E_ARG, "mh_scene_median_pass: pass %d of %d", pass, npass);
    if (pass == 0) {
        if (which == 0) {
            k_scene_ab<<<mh_cdiv(d.T, 128), 128, 0, st>>>(c->params + c->off[MH_P_ZMIN_LIN], c->params + c->off[MH_P_ZMAX_LIN], d.T, s->ab);
            MH_LAUNCHED(c);
        }
        k_med_count<<<grid, 256, 0, st>>>(s->back, d.T, HW, s->hist);
        MH_LAUNCHED(c);
        return MH_OK;
    }
    if (pass == 1) {
        k_med_start<<<grid, 256, 0, st>>>(s->hist, HW, s->ntot, s->prefix, s->krem, which == 0 ? 1 : 3);
        MH_LAUNCHED(c);
    }
    const int ndig = which == 0 ? 8 : 2;
    if (pass >= 2) {                                               // descend on the (all-reduced) histogram of the previous digit pass
        const int prev_shift = 4 * (ndig - (pass - 1));
        if (which == 0) k_med_descend<<<grid, 256, 0, st>>>(s->hist, HW, prev_shift, s->prefix, s->krem);
        else k_med_descend_img<<<grid, 256, 0, st>>>(s->hist, HW, prev_shift, s->prefix, s->krem);
        MH_LAUNCHED(c);
    }
    if (pass <= ndig) {
        const int shift = 4 * (ndig - pass);
        if (which == 0) k_med_digit_depth<<<grid, 256, 0, st>>>(c->depth, s->back, s->ab, d.T, HW, s->prefix, shift, s->hist);
        else k_med_digit_img<<<grid, 256, 0, st>>>(s->images, s->back, d.T, HW, s->prefix, shift, s->hist);
        MH_LAUNCHED(c);
    } else {
        if (which == 0) k_med_upper_depth<<<grid, 256, 0, st>>>(c->depth, s->back, s->ab, d.T, HW, s->prefix, s->aux);
        else k_med_upper_img<<<grid, 256, 0, st>>>(s->images, s->back, d.T, HW, s->prefix, s->aux);
        MH_LAUNCHED(c);
    }
    return MH_OK;
}

extern "C" int mh_scene_median_finish(mh_ctx* c, int32_t which, float* depth_host, uint8_t* mask_host, uint8_t* img_host, void* stream) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    MhSceneState* s = scene_state(c);
    if (!s) MH_FAIL(c, MH_E_STATE, "mh_scene_median_finish: no scene state");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = s->HW;
    const int grid = mh_cdiv(HW, 256);
    if (which == 0) {
        k_med_finish_depth<<<grid, 256, 0, st>>>(s->aux, s->prefix, s->ntot, HW, s->out_depth, s->out_mask);
        MH_LAUNCHED(c);
        if (depth_host) MH_CUDA(c, cudaMemcpyAsync(depth_host, s->out_depth, sizeof(float) * HW, cudaMemcpyDeviceToHost, st));
        if (mask_host) MH_CUDA(c, cudaMemcpyAsync(mask_host, s->out_mask, HW, cudaMemcpyDeviceToHost, st));
    } else {
        k_med_finish_img<<<grid, 256, 0, st>>>(s->aux, s->prefix, s->ntot, HW, s->out_img);
        MH_LAUNCHED(c);
        if (img_host) MH_CUDA(c, cudaMemcpyAsync(img_host, s->out_img, 3 * HW, cudaMemcpyDeviceToHost, st));
    }
    MH_CUDA(c, cudaStreamSynchronize(st));
    return MH_OK;
}

// Device-resident scene update of one fit() cycle (optimizer.py:578-584) after the median passes: median depth -> postprocess_depthmap
// (mh_scenepost.cu) -> scene point cloud; nothing but the hole counts of the fill-in sweeps crosses the bus.  depth_host_or_null receives
// the post-processed depth map (get_optimized_variables()['scene_depth']) when the caller wants it.
extern "C" int mh_scene_update_from_median(mh_ctx* c, int32_t use_bilateral, int32_t fillin_ksize, float* depth_host_or_null, void* stream) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    MhSceneState* s = scene_state(c);
    if (!s) MH_FAIL(c, MH_E_STATE, "mh_scene_update_from_median: no scene state (mh_scene_set_back + median passes first)");
    if (!c->camera_set) MH_FAIL(c, MH_E_STATE, "mh_scene_update_from_median: mh_set_camera first");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = s->HW;
    k_med_finish_depth<<<mh_cdiv(HW, 256), 256, 0, st>>>(s->aux, s->prefix, s->ntot, HW, s->out_depth, s->out_mask);
    MH_LAUNCHED(c);
    MH_TRY(mh_scene_postprocess_dev(c, s->out_depth, s->out_mask, use_bilateral, fillin_ksize, st));
    const float* res = mh_scene_post_result(c);
    MH_TRY(mh_scene_from_depth(c, res, s->out_mask, st));
    if (depth_host_or_null) {
        MH_CUDA(c, cudaMemcpyAsync(depth_host_or_null, res, sizeof(float) * HW, cudaMemcpyDeviceToHost, st));
        MH_CUDA(c, cudaStreamSynchronize(st));
    }
    return MH_OK;
}

// postprocess_depthmap (utils.py:174-209) of a HOST depth map on the device (testing aid and stand-alone use): mask_host_or_null u8 {0,1}
extern "C" int mh_postprocess_depthmap(mh_ctx* c, const float* depth_host, const uint8_t* mask_host_or_null, int32_t use_bilateral,
                                       int32_t fillin_ksize, float* out_host, void* stream) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    if (!depth_host || !out_host) MH_FAIL(c, MH_E_ARG, "mh_postprocess_depthmap: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)c->d.H * c->d.W;
    float* dd = nullptr; uint8_t* dm = nullptr;
    MH_CUDA(c, mh_dev_alloc((void**)&dd, HW * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(dd, depth_host, HW * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && mask_host_or_null) {
        e = mh_dev_alloc((void**)&dm, HW);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dm, mask_host_or_null, HW, cudaMemcpyHostToDevice, st);
    }
    int r = MH_OK;
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "mh_postprocess_depthmap: %s", cudaGetErrorString(e)); r = MH_E_CUDA; }
    if (r == MH_OK) r = mh_scene_postprocess_dev(c, dd, dm, use_bilateral, fillin_ksize, st);
    if (r == MH_OK && cudaMemcpyAsync(out_host, mh_scene_post_result(c), HW * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) r = MH_E_CUDA;
    cudaStreamSynchronize(st);
    mh_dev_free(dd);
    if (dm) mh_dev_free(dm);
    return r;
}

int mh_scene_views(mh_ctx* c, int which, void** ptr, int64_t* n) {
    MhSceneState* s = scene_state(c);
    if (!s) MH_FAIL(c, MH_E_STATE, "scene median buffers do not exist yet (mh_scene_set_back first)");
    if (which == MH_BUF_MEDIAN_HIST) { *ptr = s->hist; *n = 16 * s->HW * 3; }
    else { *ptr = s->aux; *n = 6 * s->HW; }
    return MH_OK;
}
