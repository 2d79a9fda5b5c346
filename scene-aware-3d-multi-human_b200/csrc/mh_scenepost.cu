// Post-processing of the aggregated scene depth map on the device (remainder): what fit() runs on the ONE
// median depth map every cycle >= 30 before it becomes the scene point cloud.
//
// Reference code replaced: postprocess_depthmap (mhmocap/utils.py:174-209) -- cv2.bilateralFilter(1 / clip(depth, .01, 100), 9,
// sigmaColor .05, sigmaSpace 25), Sobel gradients of disparity and depth, the edge threshold 3 x mean of their std-normalised sum,
// two 3x3 erosions of the keep-mask -- and the Python-loop fillin_values (utils.py:91-135: 7x7 masked median, repeated until no hole
// is left), called from optimizer.py:583-584.
//
// OpenCV semantics followed (float32 single-channel path of bilateralFilter, cv::Sobel ksize 3, cv::erode 3x3):
//   * borders: BORDER_REFLECT_101 for the bilateral filter and Sobel; erosion treats the outside as "keep";
//   * bilateral: circular support r <= 4, space weights exp(-r^2 / (2 sigmaSpace^2)); colour weights from a 4096-bin table of
//     exp(-d^2 / (2 sigmaColor^2)) over [0, max - min] with linear interpolation; weights and sums accumulated in float32 in
//     row-major order of the support; an image with max - min < FLT_EPSILON is copied;
//   * np.std / np.mean of the gradient maps: accumulated here in float64 over fixed-order partials (numpy: float32 pairwise), so
//     a pixel within ~1e-6 (relative) of the edge threshold may fall on the other side -- tests/test_gpu_scene.py bounds this.
// A hole pixel of the fill-in only reads pixels that were valid BEFORE the sweep (validity comes from the input mask and valid
// pixels are never rewritten), so one sweep is order-independent: one thread per hole pixel, ping-pong buffers.
#include "mh_ctx.h"

#include <cmath>

#define SP_LUT 4096
#define SP_MAXOFF 81
#define SP_SWEEPS 4          // fill-in sweeps enqueued per host look at the hole counters
#define SP_NPART 256

struct MhScenePost {
    float* a; float* b; float* g1; float* g2;        // (HW) work planes
    float* dep[2]; uint8_t* msk[2];                  // ping-pong of the fill-in
    float* lut;                                      // SP_LUT + 2 colour weights, then [scale_index, degenerate flag]
    double* part;                                    // (4, SP_NPART) partial sums
    double* stat;                                    // [min, max, mean_a, mean_b, std_a, std_b, mean_g]
    int* holes;                                      // [holes left, holes filled in the last sweep]
    float* offw; int* offdy; int* offdx; int noff;   // bilateral support
    int64_t HW;
    int result;                                      // ping-pong index holding the finished depth
    bool ready;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__global__ void k_sp_recip(const float* __restrict__ in, float lo, float hi, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = 1.0f / fminf(fmaxf(in[i], lo), hi);
}

// fixed-order partials (SP_NPART blocks x a fixed tree inside the block): min / max, or sum of f(x)
template <int MODE>      // 0: min & max ; 1: sum x ; 2: sum (x - mean)^2 with mean = stat[mslot]
__global__ void __launch_bounds__(256) k_sp_reduce(const float* __restrict__ x, int64_t n, const double* __restrict__ stat, int mslot,
                                                   double* __restrict__ part0, double* __restrict__ part1) {
    __shared__ double s0[256], s1[256];
    const int tid = threadIdx.x;
    double a0 = MODE == 0 ? (double)INFINITY : 0.0, a1 = MODE == 0 ? -(double)INFINITY : 0.0;
    const double mean = MODE == 2 ? stat[mslot] : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + tid; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = (double)x[i];
        if (MODE == 0) { a0 = fmin(a0, v); a1 = fmax(a1, v); }
        else if (MODE == 1) a0 += v;
        else a0 += (v - mean) * (v - mean);
    }
    s0[tid] = a0; s1[tid] = a1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            if (MODE == 0) { s0[tid] = fmin(s0[tid], s0[tid + o]); s1[tid] = fmax(s1[tid], s1[tid + o]); }
            else s0[tid] += s0[tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) { part0[blockIdx.x] = s0[0]; if (MODE == 0) part1[blockIdx.x] = s1[0]; }
}

// final step of a reduction: stat[slot] = min | max | sum / n | sqrt(sum / n)
__global__ void k_sp_final(const double* __restrict__ part0, const double* __restrict__ part1, int nblk, int mode, double n, double* __restrict__ stat,
                           int slot) {
    if (threadIdx.x != 0) return;
    if (mode == 0) {
        double mn = INFINITY, mx = -INFINITY;
        for (int i = 0; i < nblk; ++i) { mn = fmin(mn, part0[i]); mx = fmax(mx, part1[i]); }
        stat[slot] = mn; stat[slot + 1] = mx;
    } else {
        double a = 0.0;
        for (int i = 0; i < nblk; ++i) a += part0[i];
        stat[slot] = mode == 1 ? a / n : sqrt(a / n);
    }
}

// colour-weight table of cv::bilateralFilter (32f): bins over [0, max - min], exp in double, stored as float
__global__ void k_sp_lut(const double* __restrict__ stat, float sigma_color, float* __restrict__ lut) {
    const float mn = (float)stat[0], mx = (float)stat[1];
    const float len = mx - mn;
    const bool degenerate = fabsf(mn - mx) < 1.1920929e-07f;
    const float scale_index = (float)SP_LUT / len;
    const double gc = -0.5 / ((double)sigma_color * (double)sigma_color);
    for (int i = threadIdx.x; i < SP_LUT + 2; i += blockDim.x) {
        const double val = (double)i / (double)scale_index;
        lut[i] = (float)exp(val * val * gc);
    }
    if (threadIdx.x == 0) { lut[SP_LUT + 2] = scale_index; lut[SP_LUT + 3] = degenerate ? 1.0f : 0.0f; }
}

__global__ void __launch_bounds__(256) k_sp_bilateral(const float* __restrict__ src, int H, int W, const float* __restrict__ lut,
                                                      const float* __restrict__ offw, const int* __restrict__ offdy, const int* __restrict__ offdx,
                                                      int noff, float* __restrict__ dst) {
    __shared__ float sw[SP_MAXOFF];
    __shared__ int sdy[SP_MAXOFF], sdx[SP_MAXOFF];
    for (int k = threadIdx.x; k < noff; k += blockDim.x) { sw[k] = offw[k]; sdy[k] = offdy[k]; sdx[k] = offdx[k]; }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const float v0 = src[(size_t)y * W + x];
    if (lut[SP_LUT + 3] != 0.f) { dst[(size_t)y * W + x] = v0; return; }
    const float scale_index = lut[SP_LUT + 2];
    float sum = 0.f, wsum = 0.f;
    for (int k = 0; k < noff; ++k) {
        const float v = src[(size_t)reflect101(y + sdy[k], H) * W + reflect101(x + sdx[k], W)];
        float alpha = __fmul_rn(fabsf(v - v0), scale_index);
        const int idx = (int)floorf(alpha);
        alpha -= (float)idx;
        const float e0 = lut[idx], e1 = lut[idx + 1];
        const float w = __fmul_rn(sw[k], __fadd_rn(e0, __fmul_rn(alpha, e1 - e0)));
        sum = __fadd_rn(sum, __fmul_rn(v, w));
        wsum = __fadd_rn(wsum, w);
    }
    dst[(size_t)y * W + x] = sum / wsum;
}

// |Sobel_x| + |Sobel_y|, ksize 3, BORDER_REFLECT_101
__global__ void __launch_bounds__(256) k_sp_sobel(const float* __restrict__ src, int H, int W, float* __restrict__ dst) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    float p[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) p[dy + 1][dx + 1] = src[(size_t)reflect101(y + dy, H) * W + reflect101(x + dx, W)];
    const float gx = (p[0][2] - p[0][0]) + 2.0f * (p[1][2] - p[1][0]) + (p[2][2] - p[2][0]);
    const float gy = (p[2][0] - p[0][0]) + 2.0f * (p[2][1] - p[0][1]) + (p[2][2] - p[0][2]);
    dst[(size_t)y * W + x] = fabsf(gx) + fabsf(gy);
}

// g = g_disp / std_disp + g_depth / std_depth (in place of g_disp)
__global__ void k_sp_combine(float* __restrict__ g1, const float* __restrict__ g2, const double* __restrict__ stat, int64_t n) {
    const float s1 = (float)stat[4], s2 = (float)stat[5];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g1[i] = g1[i] / s1 + g2[i] / s2;
}

// keep = 1 - (g > 3 mean(g))
__global__ void k_sp_edges(const float* __restrict__ g, const double* __restrict__ stat, int64_t n, uint8_t* __restrict__ keep) {
    const float thr = 3.0f * (float)stat[6];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) keep[i] = g[i] > thr ? 0 : 1;
}

// 3x3 erosion (outside counts as keep); `mask` (or null) is multiplied into the result of the LAST erosion
__global__ void __launch_bounds__(256) k_sp_erode(const uint8_t* __restrict__ in, int H, int W, const uint8_t* __restrict__ mask, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    uint8_t v = 1;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) v &= in[(size_t)yy * W + xx];
        }
    if (mask) v = (v && mask[(size_t)y * W + x]) ? 1 : 0;
    out[(size_t)y * W + x] = v;
}

// one fill-in sweep (utils.py:91-135): a hole with a valid pixel in its (2k+1)^2 window takes the median of the valid ones
__global__ void __launch_bounds__(128) k_sp_fill(const float* __restrict__ din, const uint8_t* __restrict__ min_, int H, int W, int k,
                                                 float* __restrict__ dout, uint8_t* __restrict__ mout, int* __restrict__ holes) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    if (min_[i]) { dout[i] = din[i]; mout[i] = 1; return; }
    float v[121];                                       // up to 11 x 11 (the image fill-in of fit() uses 11)
    int n = 0;
    for (int yy = max(0, y - k); yy < min(H, y + k + 1); ++yy)
        for (int xx = max(0, x - k); xx < min(W, x + k + 1); ++xx)
            if (min_[(size_t)yy * W + xx]) {
                const float a = din[(size_t)yy * W + xx];
                int j = n++;
                while (j > 0 && v[j - 1] > a) { v[j] = v[j - 1]; --j; }      // insertion sort
                v[j] = a;
            }
    if (n == 0) { dout[i] = din[i]; mout[i] = 0; atomicAdd(holes, 1); return; }
    dout[i] = (n & 1) ? v[n / 2] : (float)(((double)v[n / 2 - 1] + (double)v[n / 2]) * 0.5);       // np.median: mean of the two middle values
    mout[i] = 1;
    atomicAdd(holes + 1, 1);
}

// -------------------------------------------------------------------------------------------------
static MhScenePost* post_state(mh_ctx* c, int64_t HW) {
    if (c->scene_post) return reinterpret_cast<MhScenePost*>(c->scene_post);
    MhScenePost* s = new MhScenePost();
    memset(s, 0, sizeof(*s));
    s->HW = HW;
    cudaError_t e = cudaSuccess;
    float** planes[] = {&s->a, &s->b, &s->g1, &s->g2, &s->dep[0], &s->dep[1]};
    for (float** p : planes) if (e == cudaSuccess) e = mh_dev_alloc((void**)p, sizeof(float) * HW);
    for (int i = 0; i < 2; ++i) if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->msk[i], HW);
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->lut, sizeof(float) * (SP_LUT + 4));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->part, sizeof(double) * 4 * SP_NPART);
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->stat, sizeof(double) * 8);
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->holes, sizeof(int) * 2 * SP_SWEEPS);
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->offw, sizeof(float) * SP_MAXOFF);
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->offdy, sizeof(int) * SP_MAXOFF);
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&s->offdx, sizeof(int) * SP_MAXOFF);
    if (e != cudaSuccess) { delete s; return nullptr; }
    c->scene_post = s;
    return s;
}

void mh_scenepost_free(mh_ctx* c) {
    MhScenePost* s = reinterpret_cast<MhScenePost*>(c->scene_post);
    if (!s) return;
    void* ptrs[] = {s->a, s->b, s->g1, s->g2, s->dep[0], s->dep[1], s->msk[0], s->msk[1], s->lut, s->part, s->stat, s->holes, s->offw, s->offdy, s->offdx};
    for (void* p : ptrs) if (p) mh_dev_free(p);
    delete s;
    c->scene_post = nullptr;
}

static int reduce_to(mh_ctx* c, MhScenePost* s, const float* x, int mode, int mslot, int slot, cudaStream_t st) {
    const int nblk = (int)std::min<int64_t>(SP_NPART, (s->HW + 255) / 256);
    if (mode == 0) k_sp_reduce<0><<<nblk, 256, 0, st>>>(x, s->HW, s->stat, 0, s->part, s->part + SP_NPART);
    else if (mode == 1) k_sp_reduce<1><<<nblk, 256, 0, st>>>(x, s->HW, s->stat, 0, s->part, s->part + SP_NPART);
    else k_sp_reduce<2><<<nblk, 256, 0, st>>>(x, s->HW, s->stat, mslot, s->part, s->part + SP_NPART);
    MH_LAUNCHED(c);
    k_sp_final<<<1, 32, 0, st>>>(s->part, s->part + SP_NPART, nblk, mode, (double)s->HW, s->stat, slot);
    MH_LAUNCHED(c);
    return MH_OK;
}

// depth_dev (H, W) f32 and mask_dev (H, W) u8 {0,1} or null -> post-processed depth in the state (mh_scene_post_result)
int mh_scene_postprocess_dev(mh_ctx* c, const float* depth_dev, const uint8_t* mask_dev, int use_bilateral, int fillin_ksize, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int64_t HW = (int64_t)d.H * d.W;
    if (fillin_ksize < 3 || fillin_ksize > 11 || !(fillin_ksize & 1)) MH_FAIL(c, MH_E_ARG, "scene post-processing: fill-in window %d (odd, 3..11)", fillin_ksize);
    MhScenePost* s = post_state(c, HW);
    if (!s) MH_FAIL(c, MH_E_CUDA, "scene post-processing: out of device memory");
    const int g1d = (int)std::min<int64_t>((HW + 255) / 256, 4096);
    const dim3 g2d(mh_cdiv(d.W, 32), mh_cdiv(d.H, 8));
    const float* depth = depth_dev;
    if (use_bilateral) {
        if (!s->noff) {                                            // support of cv::bilateralFilter(d = 9, sigmaSpace = 25)
            const int radius = 4;
            const double gs = -0.5 / (25.0 * 25.0);
            float w[SP_MAXOFF]; int dy[SP_MAXOFF], dx[SP_MAXOFF];
            int n = 0;
            for (int i = -radius; i <= radius; ++i)
                for (int j = -radius; j <= radius; ++j) {
                    const double r = std::sqrt((double)i * i + (double)j * j);
                    if (r > radius) continue;
                    w[n] = (float)std::exp(r * r * gs); dy[n] = i; dx[n] = j; ++n;
                }
            MH_CUDA(c, cudaMemcpyAsync(s->offw, w, sizeof(float) * n, cudaMemcpyHostToDevice, st));
            MH_CUDA(c, cudaMemcpyAsync(s->offdy, dy, sizeof(int) * n, cudaMemcpyHostToDevice, st));
            MH_CUDA(c, cudaMemcpyAsync(s->offdx, dx, sizeof(int) * n, cudaMemcpyHostToDevice, st));
            MH_CUDA(c, cudaStreamSynchronize(st));                 // the host arrays are on this stack frame
            s->noff = n;
        }
        k_sp_recip<<<g1d, 256, 0, st>>>(depth_dev, 0.01f, 100.0f, HW, s->a);                    // disparity
        MH_LAUNCHED(c);
        MH_TRY(reduce_to(c, s, s->a, 0, 0, 0, st));
        k_sp_lut<<<1, 256, 0, st>>>(s->stat, 0.05f, s->lut);
        MH_LAUNCHED(c);
        k_sp_bilateral<<<g2d, 256, 0, st>>>(s->a, d.H, d.W, s->lut, s->offw, s->offdy, s->offdx, s->noff, s->b);
        MH_LAUNCHED(c);
        k_sp_recip<<<g1d, 256, 0, st>>>(s->b, 0.01f, 100.0f, HW, s->dep[0]);                     // depth = 1 / clip(disp, .01, 100)
        MH_LAUNCHED(c);
        depth = s->dep[0];
    } else {
        MH_CUDA(c, cudaMemcpyAsync(s->dep[0], depth_dev, sizeof(float) * HW, cudaMemcpyDeviceToDevice, st));
        depth = s->dep[0];
    }
    k_sp_recip<<<g1d, 256, 0, st>>>(depth, 0.1f, 100.0f, HW, s->a);                               // disp = 1 / clip(depth, .1, 100)
    MH_LAUNCHED(c);
    k_sp_sobel<<<g2d, 256, 0, st>>>(s->a, d.H, d.W, s->g1);
    MH_LAUNCHED(c);
    k_sp_sobel<<<g2d, 256, 0, st>>>(depth, d.H, d.W, s->g2);
    MH_LAUNCHED(c);
    MH_TRY(reduce_to(c, s, s->g1, 1, 0, 2, st));               // mean, then std (np.std: sqrt(mean((x - mean)^2)))
    MH_TRY(reduce_to(c, s, s->g1, 2, 2, 4, st));
    MH_TRY(reduce_to(c, s, s->g2, 1, 0, 3, st));
    MH_TRY(reduce_to(c, s, s->g2, 2, 3, 5, st));
    k_sp_combine<<<g1d, 256, 0, st>>>(s->g1, s->g2, s->stat, HW);
    MH_LAUNCHED(c);
    MH_TRY(reduce_to(c, s, s->g1, 1, 0, 6, st));
    k_sp_edges<<<g1d, 256, 0, st>>>(s->g1, s->stat, HW, s->msk[1]);
    MH_LAUNCHED(c);
    uint8_t* e1 = reinterpret_cast<uint8_t*>(s->b);            // plane b is free again: scratch of the first erosion
    k_sp_erode<<<g2d, 256, 0, st>>>(s->msk[1], d.H, d.W, nullptr, e1);
    MH_LAUNCHED(c);
    k_sp_erode<<<g2d, 256, 0, st>>>(e1, d.H, d.W, mask_dev, s->msk[0]);
    MH_LAUNCHED(c);
    // fill-in sweeps until no hole is left (or no hole can be reached any more: an all-hole image).  SP_SWEEPS sweeps are enqueued per
    // look at the hole counters: a sweep over an image without holes is a plain copy, so the sweeps past the last useful one change
    // nothing -- and the host waits for the device once per group instead of once per sweep
    int cur = 0;
    const dim3 gfill(mh_cdiv(d.W, 32), mh_cdiv(d.H, 4));
    bool done = false;
    for (int sweep = 0; sweep < d.H + d.W && !done; sweep += SP_SWEEPS) {
        MH_CUDA(c, cudaMemsetAsync(s->holes, 0, sizeof(int) * 2 * SP_SWEEPS, st));
        for (int g = 0; g < SP_SWEEPS; ++g) {
            k_sp_fill<<<gfill, 128, 0, st>>>(s->dep[cur], s->msk[cur], d.H, d.W, fillin_ksize / 2, s->dep[cur ^ 1], s->msk[cur ^ 1], s->holes + 2 * g);
            MH_LAUNCHED(c);
            cur ^= 1;
        }
        int h[2 * SP_SWEEPS];
        MH_CUDA(c, cudaMemcpyAsync(h, s->holes, sizeof(h), cudaMemcpyDeviceToHost, st));
        MH_CUDA(c, cudaStreamSynchronize(st));
        for (int g = 0; g < SP_SWEEPS && !done; ++g) {
            if (h[2 * g] == 0) done = true;
            else if (h[2 * g + 1] == 0) MH_FAIL(c, MH_E_STATE, "scene post-processing: %d pixels can never be filled in (empty keep-mask)", h[2 * g]);
        }
    }
    s->result = cur;
    s->ready = true;
    return MH_OK;
}

const float* mh_scene_post_result(mh_ctx* c) {
    MhScenePost* s = reinterpret_cast<MhScenePost*>(c->scene_post);
    return (s && s->ready) ? s->dep[s->result] : nullptr;
}
