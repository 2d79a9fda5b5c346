// One-Euro refresh of the temporal targets and the scene point cloud from a scene depth map.
//
// Reference code replaced: SMPLDepthSequenceOptimizer.one_euro_filter (mhmocap/optimizer.py:664-675) driving
// OneEuroFilter (mhmocap/one_euro_filter.py:16-53) on poses_T and on the (T,N,V,3) vertices every 25 cycles
// (optimizer.py:383-392) -- there a D2H copy, a numpy loop over T and an H2D copy; here one thread per series
// scanning the device-resident vertices.  update_scene_pointcloud (optimizer.py:605-616).
#include "mh_ctx.h"

// The arithmetic follows numpy's float32 evaluation of the reference expression by expression:
//   t_i = t_{i-1} + f32(i / frame_rate)                        (cumulative-time quirk, optimizer.py:671)
//   r = f32(2 pi d_cutoff) * t_e ; a_d = r / (r + 1) ; dx = (x - x_prev) / t_e ; dx_hat = a_d dx + (1 - a_d) dx_prev
//   cutoff = min_cutoff + beta |dx_hat| ; r = (f32(2 pi) * cutoff) * t_e ; a = r / (r + 1) ; x_hat = a x + (1 - a) x_prev
struct EuroState { float x_prev, dx_prev; };

__device__ __forceinline__ float euro_step(float x, float t_e, float two_pi_dc, float two_pi, float min_cutoff, float beta, EuroState& s) {
    const float r0 = __fmul_rn(two_pi_dc, t_e);
    const float a_d = __fdiv_rn(r0, __fadd_rn(r0, 1.0f));
    const float dx = __fdiv_rn(__fsub_rn(x, s.x_prev), t_e);
    const float dx_hat = __fadd_rn(__fmul_rn(a_d, dx), __fmul_rn(__fsub_rn(1.0f, a_d), s.dx_prev));
    const float cutoff = __fadd_rn(min_cutoff, __fmul_rn(beta, fabsf(dx_hat)));
    const float r1 = __fmul_rn(__fmul_rn(two_pi, cutoff), t_e);
    const float a = __fdiv_rn(r1, __fadd_rn(r1, 1.0f));
    const float x_hat = __fadd_rn(__fmul_rn(a, x), __fmul_rn(__fsub_rn(1.0f, a), s.x_prev));
    s.x_prev = x_hat; s.dx_prev = dx_hat;
    return x_hat;
}

// series layout: (T slots, row stride) ; one thread per element of a row
__global__ void k_one_euro(const float* __restrict__ x, float* __restrict__ y, int64_t row_elems, int64_t row_stride, int T, int t0,
                           int first, const float* __restrict__ carry_in, float* __restrict__ carry_out, int64_t carry_stride,
                           float frame_rate, float min_cutoff, float beta) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= row_elems) return;
    const float two_pi = (float)(2.0 * 3.141592653589793);
    const float two_pi_dc = (float)(2.0 * 3.141592653589793 * 1.0);
    // replay the cumulative time up to the first local frame
    float t = 0.f;
    for (int i = 1; i < t0; ++i) t = __fadd_rn(t, (float)((double)i / (double)frame_rate));
    EuroState s;
    int start = 0;
    if (first) {
        s.x_prev = x[e]; s.dx_prev = 0.f;                          // OneEuroFilter(t0 = 0, x0 = y[0], dx0 = 0)
        y[e] = x[e];
        start = 1;
    } else {
        s.x_prev = carry_in[e]; s.dx_prev = carry_in[carry_stride + e];
    }
    for (int k = start; k < T; ++k) {
        const int i = t0 + k;                                      // global frame index
        const float tn = __fadd_rn(t, (float)((double)i / (double)frame_rate));
        const float t_e = __fsub_rn(tn, t);
        t = tn;
        y[(int64_t)k * row_stride + e] = euro_step(x[(int64_t)k * row_stride + e], t_e, two_pi_dc, two_pi, min_cutoff, beta, s);
    }
    carry_out[e] = s.x_prev; carry_out[carry_stride + e] = s.dx_prev;
}

int mh_filter_run(mh_ctx* c, float mc1, float b1, float mc2, float b2, float frame_rate, int first, cudaStream_t st) {
    const mh_dims& d = c->d;
    if (!c->model_set) MH_FAIL(c, MH_E_STATE, "mh_refresh_filters: mh_set_model first");
    if (d.t0 == 0) first = 1;
    MH_TRY(mh_forward_only(c, st));                                // vertices of the CURRENT parameters (optimizer.py:385-389)
    const int64_t vrow = (int64_t)d.N * MH_LD3V, trow = (int64_t)d.N * 3;
    // vertices: slots 1..T of verts -> slots 1..T of filtered
    k_one_euro<<<mh_cdiv(vrow, 256), 256, 0, st>>>(c->verts + vrow, c->filtered + vrow, vrow, vrow, d.T, d.t0, first, c->carry_in, c->carry_out,
                                                  vrow, frame_rate, mc2, b2);
    MH_LAUNCHED(c);
    k_one_euro<<<mh_cdiv(trow, 128), 128, 0, st>>>(c->params + c->off[MH_P_POSES_T], c->transfilt, trow, trow, d.T, d.t0, first,
                                                  c->carry_in + 2 * vrow, c->carry_out + 2 * vrow, trow, frame_rate, mc1, b1);
    MH_LAUNCHED(c);
    return MH_OK;
}

// stand-alone filter of a host array (T, row_elems): SMPLDepthSequenceOptimizer.one_euro_filter (optimizer.py:664-675)
extern "C" int mh_one_euro_filter(mh_ctx* c, const float* x_host, float* y_host, int32_t T, int64_t row_elems, float min_cutoff,
                                  float beta, float frame_rate) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    if (!x_host || !y_host || T < 1 || row_elems < 1) MH_FAIL(c, MH_E_ARG, "mh_one_euro_filter: bad arguments");
    const int64_t n = (int64_t)T * row_elems;
    float *dx, *dy, *carry;
    MH_CUDA(c, mh_dev_alloc((void**)&dx, sizeof(float) * n));
    cudaError_t e = mh_dev_alloc((void**)&dy, sizeof(float) * n);
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&carry, sizeof(float) * 2 * row_elems);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x_host, sizeof(float) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_one_euro<<<mh_cdiv(row_elems, 256), 256>>>(dx, dy, row_elems, row_elems, T, 0, 1, carry, carry, row_elems, frame_rate, min_cutoff, beta);
        c->launches++;
        e = cudaMemcpy(y_host, dy, sizeof(float) * n, cudaMemcpyDeviceToHost);
    }
    mh_dev_free(dx); mh_dev_free(dy); mh_dev_free(carry);
    MH_CUDA(c, e);
    return MH_OK;
}

// -------------------------------------------------------------------------------------------------
// scene_pcd = inverse projection of the pixel centres with the scene depth, kept where mask > 0.5, in
// row-major pixel order (optimizer.py:609-615, transforms.py:128-130)
#define SC_BLOCK 1024
__global__ void k_scene_count(const uint8_t* __restrict__ mask, int64_t HW, int* __restrict__ counts) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * SC_BLOCK + threadIdx.x;
    const int on = (p < HW) && mask[p] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&s, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = s;
}

// exclusive scan of the per-block counts by ONE block: thread t owns `per` consecutive entries (a serial chain of n dependent global
// accesses by one thread was 0.2 ms at 512x512 and 1.6 ms at 1080p)
__global__ void __launch_bounds__(1024) k_scene_scan(int* __restrict__ counts, int n, int* __restrict__ total) {
    __shared__ int s[1024];
    const int per = (n + 1023) / 1024, i0 = threadIdx.x * per;
    int t = 0;
    for (int k = 0; k < per; ++k) if (i0 + k < n) t += counts[i0 + k];
    s[threadIdx.x] = t;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int u = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += u;
        __syncthreads();
    }
    int run = s[threadIdx.x] - t;
    for (int k = 0; k < per; ++k)
        if (i0 + k < n) { const int v = counts[i0 + k]; counts[i0 + k] = run; run += v; }
    if (threadIdx.x == 1023) *total = s[1023];
}

__global__ void k_scene_scatter(const float* __restrict__ depth, const uint8_t* __restrict__ mask, int W, int64_t HW,
                                const int* __restrict__ offsets, float cx, float cy, float i00, float i01, float i10, float i11,
                                int64_t cap, float* __restrict__ pcd) {
    __shared__ int wcount[SC_BLOCK / 32];
    const int64_t p = (int64_t)blockIdx.x * SC_BLOCK + threadIdx.x;
    const int on = (p < HW) && mask[p] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wcount[w] = __popc(bal);
    __syncthreads();
    if (!on) return;
    int base = offsets[blockIdx.x];
    for (int k = 0; k < w; ++k) base += wcount[k];
    const int64_t o = base + __popc(bal & ((1u << lane) - 1u));
    if (o >= cap) return;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const float du = ((float)x + 0.5f) - cx, dv = ((float)y + 0.5f) - cy;
    const float z = depth[p];
    // ptsxy = d * ((uv - c) @ inv(K[:2,:2]^T))
    pcd[3 * o] = __fmul_rn(z, __fadd_rn(__fmul_rn(du, i00), __fmul_rn(dv, i10)));
    pcd[3 * o + 1] = __fmul_rn(z, __fadd_rn(__fmul_rn(du, i01), __fmul_rn(dv, i11)));
    pcd[3 * o + 2] = z;
}

int mh_scene_from_depth(mh_ctx* c, const float* depth_dev, const uint8_t* mask_dev, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int64_t HW = (int64_t)d.H * d.W;
    const int nblk = mh_cdiv(HW, SC_BLOCK);
    if (!c->scene_counts) MH_TRY(mh_alloc_ints(c, &c->scene_counts, nblk + 1));      // kept: no allocation / free (= device synchronisation) per scene update
    int* counts = c->scene_counts;
    k_scene_count<<<nblk, SC_BLOCK, 0, st>>>(mask_dev, HW, counts);
    c->launches++;
    k_scene_scan<<<1, 1024, 0, st>>>(counts, nblk, counts + nblk);
    c->launches++;
    int total = 0;
    cudaMemcpyAsync(&total, counts + nblk, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    MH_CUDA(c, e);
    if (total > d.M_max) MH_FAIL(c, MH_E_CAPACITY, "scene cloud of %d points exceeds M_max = %lld", total, (long long)d.M_max);
    // A = K[:2,:2]^T = [[k00, k10], [k01, k11]] ; inverse in closed form
    const float a = c->K[0], b = c->K[3], cc = c->K[1], dd = c->K[4];
    const float det = a * dd - b * cc;
    const float i00 = dd / det, i01 = -b / det, i10 = -cc / det, i11 = a / det;
    k_scene_scatter<<<nblk, SC_BLOCK, 0, st>>>(depth_dev, mask_dev, d.W, HW, counts, c->K[2], c->K[5], i00, i01, i10, i11, d.M_max, c->scene);
    c->launches++;
    MH_CUDA(c, cudaGetLastError());
    if (total > 0 && total < MH_KNN) MH_FAIL(c, MH_E_ARG, "scene cloud has only %d points (< %d)", total, MH_KNN);
    c->M = total;
    return mh_knn_build(c, st);
}
