// Non-raster loss terms of one fit() cycle and of the translation-init loop, forward + analytic backward.
//
// Reference code replaced (mhmocap/optimizer.py): 2D reprojection :404-420 (+ transforms.py:57-95), pose / shape
// priors :523-526, scale priors :531-532, contact + foot sliding :483-518, velocity :560-561, filtered-vertex
// velocity :563-574, depth-range activations :683-688, translation init :740-761.  Batch-partition quirks
//  are kept: a "batch" is B consecutive GLOBAL frames.
#include "mh_ctx.h"

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum for blockDim.x <= 1024; result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* sm /*32 floats*/) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = (lane < (blockDim.x + 31) / 32) ? sm[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;
}

__device__ __forceinline__ float signf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// -------------------------------------------------------------------------------------------------
// slot-major parameter views incl. the halo frames
__global__ void k_gather(const float* __restrict__ theta, const float* __restrict__ trans, const float* __restrict__ halo, int T, int N,
                         int use_prev, int use_next, float* __restrict__ theta_all, float* __restrict__ trans_all) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)(T + 2) * N * MH_HALO;
    if (i >= total) return;
    const int e = (int)(i % MH_HALO);
    const int n = (int)((i / MH_HALO) % N);
    const int s = (int)(i / ((int64_t)MH_HALO * N));
    float v;
    if (s == 0 && use_prev) v = halo[(size_t)n * MH_HALO + e];
    else if (s == T + 1 && use_next) v = halo[(size_t)(N + n) * MH_HALO + e];
    else {
        const int t = min(max(s - 1, 0), T - 1);      // unused halo slots mirror the boundary frame (never read by a term)
        v = e < 72 ? theta[((size_t)t * N + n) * 72 + e] : trans[((size_t)t * N + n) * 3 + (e - 72)];
    }
    const size_t b = (size_t)s * N + n;
    if (e < 72) theta_all[b * 72 + e] = v; else trans_all[b * 3 + (e - 72)] = v;
}

int mh_terms_gather(mh_ctx* c, int use_prev, int use_next, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int64_t total = (int64_t)c->Ts * d.N * MH_HALO;
    k_gather<<<mh_cdiv(total, 256), 256, 0, st>>>(c->params + c->off[MH_P_POSES_SMPL], c->params + c->off[MH_P_POSES_T], c->halo_recv, d.T, d.N,
                                                 use_prev, use_next, c->theta_all, c->trans_all);
    MH_LAUNCHED(c);
    return MH_OK;
}

// -------------------------------------------------------------------------------------------------
// one warp per local person-frame: 2D reprojection term + pose prior
struct CamParams { float K[9]; float Kd[5]; int has_kd; float w17[MH_NJR]; float slack17[MH_NJR]; };

__global__ void k_body_terms(const float* __restrict__ j17, const float* __restrict__ pose2d, const float* __restrict__ theta,
                             const float* __restrict__ theta_ref, const float* __restrict__ valid, CamParams cam, int T, int N, float W,
                             float H, float thr, float coef_proj, float coef_poses, float* __restrict__ gj17,
                             float* __restrict__ g_theta, float* __restrict__ g_trans, float* __restrict__ lpart, int LP) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);       // local person-frame
    const int lane = threadIdx.x & 31;
    if (i >= T * N) return;
    const size_t b = (size_t)i + N;                                          // slot-major body
    float l2d = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (lane < MH_NJR) {
        const float* P = j17 + (b * MH_NJR + lane) * 3;
        const float* q = pose2d + ((size_t)i * MH_NJR + lane) * 3;
        const float m = (q[2] >= thr ? 1.0f : 0.0f) * cam.w17[lane];         // optimizer.py:404, 420
        float uv[2];
        const float Pv[3] = {P[0], P[1], P[2]};
        mh_project(Pv, cam.K, cam.has_kd ? cam.Kd : nullptr, uv);
        const float du = m * uv[0] / W - m * q[0] / W, dv = m * uv[1] / H - m * q[1] / H;      // optimizer.py:367-368
        l2d = du * du + dv * dv;
        float gP[3];
        mh_project_bwd(Pv, cam.K, cam.has_kd ? cam.Kd : nullptr, coef_proj * 2.0f * du * m / W, coef_proj * 2.0f * dv * m / H, gP);
        float* g = gj17 + (b * MH_NJR + lane) * 3;
        g[0] = gP[0]; g[1] = gP[1]; g[2] = gP[2];
        // J17 = R17 . V + T (1 - rowsum): the part of dL/dT that does not pass through the vertices (exactly 0 when the regressor
        // rows sum to 1, as the shipped AlphaPose / MuPoTS regressors do)
        s0 = cam.slack17[lane] * gP[0]; s1 = cam.slack17[lane] * gP[1]; s2 = cam.slack17[lane] * gP[2];
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0 && (s0 != 0.f || s1 != 0.f || s2 != 0.f)) { g_trans[(size_t)i * 3] += s0; g_trans[(size_t)i * 3 + 1] += s1; g_trans[(size_t)i * 3 + 2] += s2; }
    // pose prior: sum |valid * theta_ref - valid * theta|  (optimizer.py:523-525)
    const float vld = valid[i];
    float lp = 0.f;
    for (int e = lane; e < 72; e += 32) {
        const float df = vld * theta_ref[(size_t)i * 72 + e] - vld * theta[(size_t)i * 72 + e];
        lp += fabsf(df);
        g_theta[(size_t)i * 72 + e] = -coef_poses * vld * signf(df);
    }
    l2d = warp_sum(l2d); lp = warp_sum(lp);
    // loss partials: one slot per contributor, summed in a fixed order by k_loss_reduce (float atomics would make the logged
    // losses depend on the order the warps finish in)
    if (lane == 0) { lpart[(size_t)MH_L_POSE2D * LP + i] = l2d; lpart[(size_t)MH_L_REF_POSES * LP + i] = lp; }
}

// -------------------------------------------------------------------------------------------------
// dL/dV initialiser for the local bodies: filtered-vertex velocity term (optimizer.py:563-574) + the
// regressed-joint gradients spread through R17^T.  Writes every column of the padded row.
__global__ void __launch_bounds__(256) k_dverts_init(const float* __restrict__ verts, const float* __restrict__ filt,
                                                     const float* __restrict__ gj17, const int* __restrict__ cptr,
                                                     const int* __restrict__ cjoint, const float* __restrict__ cw, int T, int N, int t0,
                                                     int T_total, int use_prev, int use_next, int has_filters, float coef,
                                                     float* __restrict__ dverts, float* __restrict__ lpart, int LP) {
    __shared__ float sm[32];
    const int i = blockIdx.y;                    // local person-frame
    const int s = i / N + 1;                     // slot
    const size_t b = (size_t)i + N;
    const int tg = t0 + s - 1;
    const bool has_prev = has_filters && tg >= 1 && (s > 1 || use_prev);
    const bool has_next = has_filters && (tg + 1 <= T_total - 1) && (s < T || use_next);
    const size_t row = b * MH_LD3V, rp = row - (size_t)N * MH_LD3V, rn = row + (size_t)N * MH_LD3V;
    float loss = 0.f;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < MH_LD3V; e += gridDim.x * blockDim.x) {
        float g = 0.f;
        if (e < 3 * MH_V) {
            const float v = verts[row + e];
            if (has_prev) {
                const float D = (v - verts[rp + e]) - (filt[row + e] - filt[rp + e]);
                loss += D * D;
                g += 2.0f * coef * D;
            }
            if (has_next) {
                const float D = (verts[rn + e] - v) - (filt[rn + e] - filt[row + e]);
                g -= 2.0f * coef * D;
            }
            const int vtx = e / 3, k = e - 3 * vtx;
            for (int q = cptr[vtx]; q < cptr[vtx + 1]; ++q) g = fmaf(cw[q], gj17[(b * MH_NJR + cjoint[q]) * 3 + k], g);
        }
        dverts[row + e] = g;
    }
    if (has_filters) {
        loss = block_sum(loss, sm);
        if (threadIdx.x == 0) lpart[(size_t)MH_L_FILTER_VERTS * LP + (size_t)blockIdx.y * gridDim.x + blockIdx.x] = loss;
    }
}

// -------------------------------------------------------------------------------------------------
// velocity term on the translations (optimizer.py:560-561; init: :758-760)
__global__ void k_velocity(const float* __restrict__ trans_all, int T, int N, int t0, int T_total, int use_prev, int use_next, float coef,
                           float* __restrict__ g_trans, float* __restrict__ lpart, int LP) {
    __shared__ float sm[32];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;        // (t, n, k)
    float loss = 0.f;
    if (idx < T * N * 3) {
        const int s = idx / (N * 3) + 1;
        const int tg = t0 + s - 1;
        const size_t o = (size_t)idx + (size_t)N * 3;
        const float v = trans_all[o];
        float g = 0.f;
        if (tg >= 1 && (s > 1 || use_prev)) {
            const float D = v - trans_all[o - (size_t)N * 3];
            loss = D * D;
            g += 2.0f * coef * D;
        }
        if (tg + 1 <= T_total - 1 && (s < T || use_next)) g -= 2.0f * coef * (trans_all[o + (size_t)N * 3] - v);
        g_trans[idx] += g;
    }
    loss = block_sum(loss, sm);
    if (threadIdx.x == 0) lpart[(size_t)MH_L_VEL * LP + blockIdx.x] = loss;
}

// -------------------------------------------------------------------------------------------------
// Contact term: streaming exact top-32 nearest scene points of the lowest vertex (optimizer.py:487-506).
// One CTA per KNN_Q consecutive local person-frames: every scene point is loaded once and tested against all of them (the
// cloud is L2-resident; one CTA per person-frame is L2-bandwidth bound).  KNN_Q is chosen so that the grid still fills the GPU twice.  Pass A: per-thread minima -> the 32nd
// smallest of them bounds the 32nd nearest distance; pass B collects every point under the bound; the 32 nearest are ranked
// exactly by (distance, index).  No distance matrix, no sort over M.
#define KNN_THREADS 256
#define KNN_CAP 512
__device__ __forceinline__ float d2_point(float p0, float p1, float p2, float x, float y, float z) {
    // sum(pow(pcd - low, 2), -1): three squares added in coordinate order
    const float a = p0 - x, b = p1 - y, c = p2 - z;
    return (a * a + b * b) + c * c;
}

template <int KNN_Q>
__global__ void __launch_bounds__(KNN_THREADS) k_contact(const float* __restrict__ verts, const int* __restrict__ lowidx,
                                                         const float* __restrict__ scene, int64_t M, int N, int TN, float coef,
                                                         float* __restrict__ contact, float* __restrict__ g_trans,
                                                         float* __restrict__ lpart, int LP, const uint8_t* __restrict__ resolved) {
    if (resolved) {               // the grid search (k_contact_grid) answered these person-frames already: nothing to stream
        bool all = true;
        for (int q = 0; q < KNN_Q; ++q) { const int i = blockIdx.x * KNN_Q + q; all = all && (i >= TN || resolved[i]); }
        if (all) return;
    }
    __shared__ float smin[KNN_Q][KNN_THREADS];
    __shared__ float cd[KNN_Q][KNN_CAP];
    __shared__ int ci[KNN_Q][KNN_CAP];
    __shared__ int sel[KNN_Q][MH_KNN];
    __shared__ int scount[KNN_Q];
    // the bound is a (distance, point index) KEY: with more than KNN_CAP points at exactly the bounding distance (duplicated scene
    // points) a bound on the distance alone cannot be bisected; on the key it always can (all keys are distinct)
    __shared__ unsigned long long stau[KNN_Q], slo[KNN_Q], shi[KNN_Q];
    __shared__ int sdone[KNN_Q];
    const int i0 = blockIdx.x * KNN_Q, tid = threadIdx.x;
    const int nq = min(KNN_Q, TN - i0);
    float x[KNN_Q], y[KNN_Q], z[KNN_Q], mn[KNN_Q];
#pragma unroll
    for (int q = 0; q < KNN_Q; ++q) {
        const size_t b = (size_t)min(i0 + q, TN - 1) + N;
        const int li = lowidx[b];
        x[q] = verts[b * MH_LD3V + 3 * li]; y[q] = verts[b * MH_LD3V + 3 * li + 1]; z[q] = verts[b * MH_LD3V + 3 * li + 2];
        mn[q] = INFINITY;
    }
    for (int64_t p = tid; p < M; p += KNN_THREADS) {
        const float p0 = scene[3 * p], p1 = scene[3 * p + 1], p2 = scene[3 * p + 2];
#pragma unroll
        for (int q = 0; q < KNN_Q; ++q) mn[q] = fminf(mn[q], d2_point(p0, p1, p2, x[q], y[q], z[q]));
    }
#pragma unroll
    for (int q = 0; q < KNN_Q; ++q) smin[q][tid] = mn[q];
    if (tid < KNN_Q) { scount[tid] = 0; sdone[tid] = tid >= nq; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < KNN_Q; ++q) {
        int rank = 0;
        for (int j = 0; j < KNN_THREADS; ++j) rank += (smin[q][j] < mn[q]) || (smin[q][j] == mn[q] && j < tid);
        if (rank == MH_KNN - 1) { stau[q] = ((unsigned long long)__float_as_uint(mn[q]) << 32) | 0xffffffffull; slo[q] = 0ull; shi[q] = stau[q]; }
    }
    __syncthreads();
    for (int iter = 0; iter < 64; ++iter) {
        unsigned long long tau[KNN_Q];
        bool act[KNN_Q];
#pragma unroll
        for (int q = 0; q < KNN_Q; ++q) { tau[q] = stau[q]; act[q] = !sdone[q]; }
        for (int64_t p = tid; p < M; p += KNN_THREADS) {
            const float p0 = scene[3 * p], p1 = scene[3 * p + 1], p2 = scene[3 * p + 2];
#pragma unroll
            for (int q = 0; q < KNN_Q; ++q) {
                const float dd = d2_point(p0, p1, p2, x[q], y[q], z[q]);
                if (act[q] && ((((unsigned long long)__float_as_uint(dd)) << 32) | (unsigned)p) <= tau[q]) {
                    const int k = atomicAdd(&scount[q], 1);
                    if (k < KNN_CAP) { cd[q][k] = dd; ci[q][k] = (int)p; }
                }
            }
        }
        __syncthreads();
        bool all = true;
        // volatile: without it the compiler loads sdone[0] in EVERY thread before testing tid (harmless, but a read / write hazard for racecheck)
        if (tid < KNN_Q && !reinterpret_cast<volatile int*>(sdone)[tid]) {
            const int cnt = scount[tid];
            if (cnt >= MH_KNN && cnt <= KNN_CAP) sdone[tid] = 1;
            else {               // bisection on the bound (only reached with > KNN_CAP near-ties)
                if (cnt > KNN_CAP) shi[tid] = stau[tid]; else slo[tid] = stau[tid];
                stau[tid] = slo[tid] + (shi[tid] - slo[tid]) / 2;
                scount[tid] = 0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < KNN_Q; ++q) all = all && sdone[q];
        if (all) break;
    }
    for (int q = 0; q < nq; ++q) {
        const int cnt = min(scount[q], KNN_CAP);
        for (int a = tid; a < cnt; a += KNN_THREADS) {
            const float da = cd[q][a];
            const int ia = ci[q][a];
            int r = 0;
            for (int j = 0; j < cnt; ++j) r += (cd[q][j] < da) || (cd[q][j] == da && ci[q][j] < ia);
            if (r < MH_KNN) sel[q][r] = ia;
        }
    }
    __syncthreads();
    if (tid < nq && !(resolved && resolved[i0 + tid])) {
        const int q = tid, i = i0 + q;
        float my = 0.f;
        for (int r = 0; r < MH_KNN; ++r) my += scene[3 * (size_t)sel[q][r] + 1];
        my *= (1.0f / MH_KNN);
        float yq = 0.f;
#pragma unroll
        for (int k = 0; k < KNN_Q; ++k) if (k == q) yq = y[k];
        const float cdv = my - yq;                              // contact_dist_vertical (:501)
        const float r = cdv + 0.02f;                            // target.y = T.y + cdv + 0.02 (:502-503)
        lpart[(size_t)MH_L_CONTACT * LP + i] = fabsf(r);
        g_trans[(size_t)i * 3 + 1] += coef * -signf(r);         // d|T - target|/dT.y with the target detached (:504-506)
        contact[(size_t)i * 4] = cdv;
        contact[(size_t)i * 4 + 1] = cdv > -0.20f ? 1.0f : 0.0f;   // in_thr_contact_region (:509-510)
    }
}

// -------------------------------------------------------------------------------------------------
// Uniform grid over the scene cloud for the contact term: the 32 nearest points of ONE query point lie in a few cells around it, so a
// warp per person-frame visits the cells shell by shell (Chebyshev distance r = 0, 1, 2, ...) and stops when the 32nd best distance
// is below (r h)^2 -- every unvisited point is at least r h away.  Exact: same distance arithmetic (d2_point) and the same
// (distance, index) order as the streaming kernel, which stays as the fallback for queries the grid cannot answer in KG_RMAX shells
// (a person far away from every scene point).  KG_LEVELS grids with cell sizes h, 4h, 16h: a person-frame the fine grid cannot answer
// (a jump: the lowest vertex more than 6 h above the floor) is searched again on the next coarser one.  The grids are rebuilt with
// the cloud (mh_set_scene / mh_set_scene_from_depth).
#define KG_MAXCELLS (1 << 21)
static_assert(KG_MAXCELLS / 1024 + 1 <= 4096, "k_kg_scan2 scans at most 4096 block sums");
#define KG_RMAX 6
#define KG_LEVELS 3
#define KG_BUDGET 65536   // points one warp may visit on one level before it gives the person-frame up to the next level / the streaming kernel
struct MhKnnLevel {
    float* bbox;        // [6] min xyz, max xyz, then [h, 1/h], dims as ints at +8
    int* cell_start;    // KG_MAXCELLS + 1 (counts, then exclusive offsets), + block sums
    float4* gpts;       // (M_max) points sorted by cell: x, y, z, index
};
struct MhKnnGrid {
    MhKnnLevel lev[KG_LEVELS];
    int* cursor;        // KG_MAXCELLS (build scratch)
    int* pcell;         // (M_max) cell of every point (build scratch)
    float* part;        // bbox partials
    uint8_t* resolved;  // (TN)
    int64_t M_built;    // points the grids were built for (0: no grid)
};

__global__ void __launch_bounds__(256) k_kg_bbox(const float* __restrict__ pts, int64_t M, float* __restrict__ part) {
    __shared__ float s[6][256];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < M; i += (int64_t)gridDim.x * 256)
        for (int k = 0; k < 3; ++k) { const float v = pts[3 * i + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
    for (int k = 0; k < 3; ++k) { s[k][threadIdx.x] = mn[k]; s[3 + k][threadIdx.x] = mx[k]; }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 3; ++k) {
                s[k][threadIdx.x] = fminf(s[k][threadIdx.x], s[k][threadIdx.x + o]);
                s[3 + k][threadIdx.x] = fmaxf(s[3 + k][threadIdx.x], s[3 + k][threadIdx.x + o]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 6) part[blockIdx.x * 6 + threadIdx.x] = s[threadIdx.x][0];
}

// bbox -> cell size (about 48 points per occupied cell of a 2-D surface cloud, at most KG_MAXCELLS cells) and grid dimensions
__global__ void k_kg_setup(const float* __restrict__ part, int nblk, int64_t M, float* __restrict__ bbox, float scale) {
    if (threadIdx.x != 0) return;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int b = 0; b < nblk; ++b)
        for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], part[b * 6 + k]); mx[k] = fmaxf(mx[k], part[b * 6 + 3 + k]); }
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    const float area = fmaxf(ex * ey + ey * ez + ex * ez, 1e-6f);
    float h = fminf(fmaxf(sqrtf(area * 48.0f / (float)M), 0.02f), 1.0f) * scale;
    int dx, dy, dz;
    for (;;) {
        dx = (int)(ex / h) + 1; dy = (int)(ey / h) + 1; dz = (int)(ez / h) + 1;
        if ((double)dx * dy * dz <= (double)KG_MAXCELLS && dx <= 1024 && dy <= 1024 && dz <= 1024) break;
        h *= 1.26f;
    }
    for (int k = 0; k < 3; ++k) { bbox[k] = mn[k]; bbox[3 + k] = mx[k]; }
    bbox[6] = h; bbox[7] = 1.0f / h;
    int* dims = reinterpret_cast<int*>(bbox + 8);
    dims[0] = dx; dims[1] = dy; dims[2] = dz; dims[3] = dx * dy * dz;
}

__device__ __forceinline__ int kg_coord(float v, float lo, float ih, int n) { return min(max((int)((v - lo) * ih), 0), n - 1); }

__global__ void k_kg_count(const float* __restrict__ pts, int64_t M, const float* __restrict__ bbox, int* __restrict__ pcell, int* __restrict__ count) {
    const int* dims = reinterpret_cast<const int*>(bbox + 8);
    const float ih = bbox[7];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const int cx = kg_coord(pts[3 * i], bbox[0], ih, dims[0]), cy = kg_coord(pts[3 * i + 1], bbox[1], ih, dims[1]),
                  cz = kg_coord(pts[3 * i + 2], bbox[2], ih, dims[2]);
        const int c = (cz * dims[1] + cy) * dims[0] + cx;
        pcell[i] = c;
        atomicAdd(&count[c], 1);
    }
}

// exclusive scan of count[0 .. ncell] in three steps (1024 cells per block)
__global__ void __launch_bounds__(1024) k_kg_scan1(int* __restrict__ a, const float* __restrict__ bbox, int* __restrict__ bsum) {
    __shared__ int s[1024];
    const int ncell = reinterpret_cast<const int*>(bbox + 8)[3];
    if ((int)blockIdx.x * 1024 > ncell) {                                // beyond the grid of this cloud: nothing to add up
        if (threadIdx.x == 0) bsum[blockIdx.x] = 0;
        return;
    }
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const int v = (i < ncell) ? a[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int t = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i <= ncell) a[i] = s[threadIdx.x] - v;                  // exclusive inside the block (a[ncell] gets the block-local total so far)
    if (threadIdx.x == 1023) bsum[blockIdx.x] = s[1023];
}
// exclusive scan of the block sums by ONE block (nblk <= 4096: four consecutive entries per thread)
__global__ void __launch_bounds__(1024) k_kg_scan2(int* __restrict__ bsum, int nblk) {
    __shared__ int s[1024];
    const int i0 = threadIdx.x * 4;
    int v[4], t = 0;
    for (int k = 0; k < 4; ++k) { v[k] = (i0 + k < nblk) ? bsum[i0 + k] : 0; t += v[k]; }
    s[threadIdx.x] = t;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int u = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += u;
        __syncthreads();
    }
    int run = s[threadIdx.x] - t;
    for (int k = 0; k < 4; ++k) { if (i0 + k < nblk) bsum[i0 + k] = run; run += v[k]; }
}
__global__ void __launch_bounds__(1024) k_kg_scan3(int* __restrict__ a, const float* __restrict__ bbox, const int* __restrict__ bsum, int* __restrict__ cursor) {
    const int ncell = reinterpret_cast<const int*>(bbox + 8)[3];
    const int i = blockIdx.x * 1024 + threadIdx.x;
    if (i <= ncell) { const int v = a[i] + bsum[blockIdx.x]; a[i] = v; if (i < ncell) cursor[i] = v; }
}

__global__ void k_kg_scatter(const float* __restrict__ pts, int64_t M, const int* __restrict__ pcell, int* __restrict__ cursor, float4* __restrict__ gpts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const int pos = atomicAdd(&cursor[pcell[i]], 1);
        gpts[pos] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float((int)i));
    }
}

// one WARP per local person-frame: lane l holds the (l+1)-th best (distance, index) key found so far
__global__ void __launch_bounds__(128) k_contact_grid(const float* __restrict__ verts, const int* __restrict__ lowidx, const float* __restrict__ scene,
                                                      const float4* __restrict__ gpts, const int* __restrict__ cell_start, const float* __restrict__ bbox,
                                                      int N, int TN, float coef, float* __restrict__ contact, float* __restrict__ g_trans,
                                                      float* __restrict__ lpart, int LP, uint8_t* __restrict__ resolved, int first) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= TN) return;
    if (!first && resolved[i]) return;                             // a finer grid answered it
    const int lane = threadIdx.x & 31;
    const size_t b = (size_t)i + N;
    const int li = lowidx[b];
    const float qx = verts[b * MH_LD3V + 3 * li], qy = verts[b * MH_LD3V + 3 * li + 1], qz = verts[b * MH_LD3V + 3 * li + 2];
    const int* dims = reinterpret_cast<const int*>(bbox + 8);
    const int DX = dims[0], DY = dims[1], DZ = dims[2];
    const float h = bbox[6], ih = bbox[7];
    const int cx = kg_coord(qx, bbox[0], ih, DX), cy = kg_coord(qy, bbox[1], ih, DY), cz = kg_coord(qz, bbox[2], ih, DZ);
    unsigned long long best = 0xffffffffffffffffull;              // lane l: (l+1)-th smallest key, ascending over the lanes
    bool done = false;
    int visited = 0;
    for (int r = 0; r <= KG_RMAX && !done && visited <= KG_BUDGET; ++r) {
        for (int z = cz - r; z <= cz + r; ++z) {
            if (z < 0 || z >= DZ) continue;
            for (int y = cy - r; y <= cy + r; ++y) {
                if (y < 0 || y >= DY) continue;
                const bool face = (z == cz - r) || (z == cz + r) || (y == cy - r) || (y == cy + r);
                // on the faces of the shell every x in [cx - r, cx + r] belongs to it, elsewhere only the two ends
                const int xstep = (face || r == 0) ? 1 : 2 * r;
                for (int x = cx - r; x <= cx + r; x += xstep) {
                    if (x < 0 || x >= DX) continue;
                    const int c = (z * DY + y) * DX + x;
                    const int p0 = cell_start[c], p1 = cell_start[c + 1];
                    visited += p1 - p0;
                    for (int p = p0 + lane; p - lane < p1; p += 32) {
                        unsigned long long cand = 0xffffffffffffffffull;
                        if (p < p1) {
                            const float4 g = gpts[p];
                            cand = ((unsigned long long)__float_as_uint(d2_point(g.x, g.y, g.z, qx, qy, qz)) << 32) | (unsigned)__float_as_int(g.w);
                        }
                        // insert, lowest lane first, every candidate that beats the current 32nd best
                        unsigned long long worst = __shfl_sync(0xffffffffu, best, 31);
                        unsigned todo = __ballot_sync(0xffffffffu, cand < worst);
                        while (todo) {
                            const int src = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const unsigned long long v = __shfl_sync(0xffffffffu, cand, src);
                            if (v < worst) {
                                const int pos = __popc(__ballot_sync(0xffffffffu, best < v));       // sorted position of v
                                const unsigned long long up = __shfl_up_sync(0xffffffffu, best, 1);
                                if (lane == pos) best = v; else if (lane > pos) best = up;
                                worst = __shfl_sync(0xffffffffu, best, 31);
                            }
                        }
                    }
                }
            }
        }
        // every point not visited yet is at least r h away from the query
        // (the cell of a point is computed in float: 1e-3 cells of slack cover the rounding of (v - lo) / h)
        const float rb = fmaxf((float)r - 1e-3f, 0.f) * h;
        const unsigned long long worst = __shfl_sync(0xffffffffu, best, 31);
        done = (worst != 0xffffffffffffffffull) && (__uint_as_float((unsigned)(worst >> 32)) < rb * rb);
    }
    if (lane == 0) resolved[i] = done ? 1 : 0;
    if (!done) return;                                             // the streaming kernel answers this one
    // mean height of the 32 neighbours, summed in rank order as the streaming kernel does (bit-identical result)
    float my = 0.f;
    const float yk = scene[3 * (size_t)(unsigned)(best & 0xffffffffull) + 1];
    for (int k = 0; k < MH_KNN; ++k) my += __shfl_sync(0xffffffffu, yk, k);
    if (lane == 0) {
        my *= (1.0f / MH_KNN);
        const float cdv = my - qy;                              // contact_dist_vertical (:501)
        const float r = cdv + 0.02f;                            // target.y = T.y + cdv + 0.02 (:502-503)
        lpart[(size_t)MH_L_CONTACT * LP + i] = fabsf(r);
        g_trans[(size_t)i * 3 + 1] += coef * -signf(r);         // d|T - target|/dT.y with the target detached (:504-506)
        contact[(size_t)i * 4] = cdv;
        contact[(size_t)i * 4 + 1] = cdv > -0.20f ? 1.0f : 0.0f;   // in_thr_contact_region (:509-510)
    }
}

static MhKnnGrid* knn_grid(mh_ctx* c) { return reinterpret_cast<MhKnnGrid*>(c->knn); }

void mh_knn_free(mh_ctx* c) {
    MhKnnGrid* g = knn_grid(c);
    if (!g) return;
    for (int l = 0; l < KG_LEVELS; ++l) { mh_dev_free(g->lev[l].bbox); mh_dev_free(g->lev[l].cell_start); mh_dev_free(g->lev[l].gpts); }
    mh_dev_free(g->cursor); mh_dev_free(g->pcell); mh_dev_free(g->part); mh_dev_free(g->resolved);
    delete g;
    c->knn = nullptr;
}

// testing aid: out[0] = person-frames the grid search answered in the last cycle, out[1] = cells, out[2] = points the grid holds,
// out[3] = cell size in micrometres (all 0 without a grid)
extern "C" int mh_debug_knn_stats(mh_ctx* c, int64_t* out, void* stream) {
    if (!c || !out) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    out[0] = out[1] = out[2] = out[3] = 0;
    MhKnnGrid* g = knn_grid(c);
    if (!g || g->M_built != c->M || c->M == 0) return MH_OK;
    const int TN = c->d.T * c->d.N;
    std::vector<uint8_t> r(TN);
    float bb[12];
    MH_CUDA(c, cudaMemcpyAsync(r.data(), g->resolved, TN, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MH_CUDA(c, cudaMemcpyAsync(bb, g->lev[0].bbox, sizeof(bb), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MH_CUDA(c, cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < TN; ++i) out[0] += r[i];
    out[1] = reinterpret_cast<const int*>(bb + 8)[3];
    out[2] = g->M_built;
    out[3] = (int64_t)(bb[6] * 1e6f);
    return MH_OK;
}

// (re)build the grid for the current cloud c->scene[0 .. c->M); MH_KNN_GRID=0 keeps the streaming kernel alone
int mh_knn_build(mh_ctx* c, cudaStream_t st) {
    const char* env = getenv("MH_KNN_GRID");
    const int use_grid = env ? atoi(env) : 1;
    MhKnnGrid* g = knn_grid(c);
    if (!use_grid || c->M < MH_KNN) { if (g) g->M_built = 0; return MH_OK; }
    const mh_dims& d = c->d;
    if (!g) {
        g = new MhKnnGrid();
        memset(g, 0, sizeof(*g));
        const int64_t Mmax = std::max<int64_t>(d.M_max, 1);
        cudaError_t e = mh_dev_alloc((void**)&g->cursor, sizeof(int) * KG_MAXCELLS);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&g->pcell, sizeof(int) * Mmax);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&g->part, sizeof(float) * 6 * 256);
        if (e == cudaSuccess) e = mh_dev_alloc((void**)&g->resolved, (size_t)d.T * d.N);
        for (int l = 0; l < KG_LEVELS && e == cudaSuccess; ++l) {
            e = mh_dev_alloc((void**)&g->lev[l].bbox, sizeof(float) * 16);
            if (e == cudaSuccess) e = mh_dev_alloc((void**)&g->lev[l].cell_start, sizeof(int) * (KG_MAXCELLS + 1 + 4096));
            if (e == cudaSuccess) e = mh_dev_alloc((void**)&g->lev[l].gpts, sizeof(float4) * Mmax);
        }
        c->knn = g;
        if (e != cudaSuccess) { mh_knn_free(c); MH_FAIL(c, MH_E_CUDA, "contact grid: %s", cudaGetErrorString(e)); }
    }
    const int64_t M = c->M;
    const int nb = (int)std::min<int64_t>(256, (M + 255) / 256);
    const int gp = (int)std::min<int64_t>(2048, (M + 255) / 256);
    const int nsb = KG_MAXCELLS / 1024 + 1;                        // covers index ncell <= KG_MAXCELLS
    k_kg_bbox<<<nb, 256, 0, st>>>(c->scene, M, g->part);
    MH_LAUNCHED(c);
    float scale = 1.0f;
    for (int l = 0; l < KG_LEVELS; ++l, scale *= 4.0f) {
        MhKnnLevel& v = g->lev[l];
        int* bsum = v.cell_start + KG_MAXCELLS + 1;
        k_kg_setup<<<1, 32, 0, st>>>(g->part, nb, M, v.bbox, scale);
        MH_LAUNCHED(c);
        MH_CUDA(c, cudaMemsetAsync(v.cell_start, 0, sizeof(int) * (KG_MAXCELLS + 1), st));
        k_kg_count<<<gp, 256, 0, st>>>(c->scene, M, v.bbox, g->pcell, v.cell_start);
        MH_LAUNCHED(c);
        k_kg_scan1<<<nsb, 1024, 0, st>>>(v.cell_start, v.bbox, bsum);
        MH_LAUNCHED(c);
        k_kg_scan2<<<1, 1024, 0, st>>>(bsum, nsb);
        MH_LAUNCHED(c);
        k_kg_scan3<<<nsb, 1024, 0, st>>>(v.cell_start, v.bbox, bsum, g->cursor);
        MH_LAUNCHED(c);
        k_kg_scatter<<<gp, 256, 0, st>>>(c->scene, M, g->pcell, g->cursor, v.gpts);
        MH_LAUNCHED(c);
    }
    g->M_built = M;
    return MH_OK;
}

// foot sliding (optimizer.py:512-518): one CTA per local batch segment; pairs are ADJACENT ENTRIES OF ONE BATCH.  Gather form: the
// thread of body (t, n) applies BOTH gradients that reach the body -- as the later frame of pair (t-1, t), at its own lowest vertex, and
// as the earlier frame of pair (t, t+1), at the lowest vertex of frame t+1 -- so no two threads touch the same row (no atomics: the
// sum does not depend on an execution order)
__global__ void k_foot(const float* __restrict__ verts, const int* __restrict__ lowidx, const float* __restrict__ contact, int T, int N,
                       int B, float coef, float* __restrict__ dverts, float* __restrict__ lpart, int LP) {
    __shared__ float sm[32];
    __shared__ float sden;
    const int k = blockIdx.x;
    const int ta = k * B, tb = min(ta + B, T);
    const int npairs = (tb - ta - 1) * N;
    float cnt = 0.f;
    for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
        const int t = ta + 1 + p / N, n = p % N;
        cnt += contact[((size_t)t * N + n) * 4 + 1];
    }
    cnt = block_sum(cnt, sm);
    if (threadIdx.x == 0) sden = fmaxf(cnt, 1.0f);               // clamp(sum(in), 1)
    __syncthreads();
    const float den = sden;
    float loss = 0.f;
    // gradient of pair (t - 1, t) w.r.t. coordinate q of vertex lowidx[t] of frame t (and minus that for frame t - 1)
    auto pair_grad = [&](int t, int n, int q, float* absdf) {
        const float in = contact[((size_t)t * N + n) * 4 + 1];
        const size_t bt = (size_t)(t + 1) * N + n, bp = bt - N;
        const int li = lowidx[bt];
        const float df = in * verts[bt * MH_LD3V + 3 * li + q] - in * verts[bp * MH_LD3V + 3 * li + q];
        *absdf = fabsf(df);
        return coef * in * signf(df) / den;
    };
    const int nbod = (tb - ta) * N;
    for (int p = threadIdx.x; p < nbod; p += blockDim.x) {
        const int t = ta + p / N, n = p % N;
        const size_t bt = (size_t)(t + 1) * N + n;
        float* row = dverts + bt * MH_LD3V;
        for (int q = 0; q < 3; ++q) {
            float a;
            if (t > ta) {                                         // later frame of pair (t - 1, t): counted here for the loss
                const float g = pair_grad(t, n, q, &a);
                loss += a;
                if (g != 0.f) row[3 * lowidx[bt] + q] += g;
            }
            if (t + 1 < tb) {                                     // earlier frame of pair (t, t + 1)
                const float g = pair_grad(t + 1, n, q, &a);
                if (g != 0.f) row[3 * lowidx[bt + N] + q] -= g;
            }
        }
    }
    loss = block_sum(loss, sm);
    if (threadIdx.x == 0) lpart[(size_t)MH_L_FOOT * LP + blockIdx.x] = loss / den;
}

static CamParams cam_of(const mh_ctx* c) {
    CamParams p;
    memcpy(p.K, c->K, sizeof(p.K));
    memcpy(p.Kd, c->Kd, sizeof(p.Kd));
    p.has_kd = c->has_kd ? 1 : 0;
    memcpy(p.w17, c->w17, sizeof(p.w17));
    memcpy(p.slack17, c->r17_slack, sizeof(p.slack17));
    return p;
}

int mh_terms_pre_raster(mh_ctx* c, int use_prev, int use_next, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int TN = d.T * d.N;
    float* losses = c->grads + c->n_params;
    float* g_trans = c->grads + c->off[MH_P_POSES_T];
    k_body_terms<<<mh_cdiv(TN, 4), 128, 0, st>>>(c->j17, c->pose2d, c->params + c->off[MH_P_POSES_SMPL], c->theta_ref, c->valid, cam_of(c),
                                                 d.T, d.N, (float)d.W, (float)d.H, c->c.joint_confidence_thr, c->c.proj2d, c->c.reg_poses,
                                                 c->gj17, c->grads + c->off[MH_P_POSES_SMPL], g_trans, c->lpart, c->LP);
    MH_LAUNCHED(c);
    k_dverts_init<<<dim3(8, TN), 256, 0, st>>>(c->verts, c->filtered, c->gj17, c->cptr, c->cjoint, c->cw, d.T, d.N, d.t0, d.T_total,
                                               use_prev, use_next, c->has_filters ? 1 : 0, c->c.reg_verts_filter, c->dverts, c->lpart, c->LP);
    MH_LAUNCHED(c);
    k_velocity<<<mh_cdiv(TN * 3, 256), 256, 0, st>>>(c->trans_all, d.T, d.N, d.t0, d.T_total, use_prev, use_next, c->c.reg_velocity,
                                                     g_trans, c->lpart, c->LP);
    MH_LAUNCHED(c);
    if (c->M > 0) {
        // person-frames per CTA: as many as keep two waves of CTAs -- measured: 256 CTAs of 2 are slower than 512 CTAs of 1 on 148 SMs
        // (MH_KNN_Q forces a value: testing aid)
        const int w2 = 2 * c->num_sms;
        int Q = TN >= 8 * w2 ? 8 : TN >= 4 * w2 ? 4 : TN >= 2 * w2 ? 2 : 1;
        if (const char* v = getenv("MH_KNN_Q")) { const int q = atoi(v); if (q == 1 || q == 2 || q == 4 || q == 8) Q = q; }
        // grid search first (a warp per person-frame); the streaming kernel then only runs for the person-frames it could not answer
        MhKnnGrid* kg = knn_grid(c);
        const uint8_t* resolved = nullptr;
        if (kg && kg->M_built == c->M) {
            for (int l = 0; l < KG_LEVELS; ++l) {
                k_contact_grid<<<mh_cdiv(TN, 4), 128, 0, st>>>(c->verts, c->lowidx, c->scene, kg->lev[l].gpts, kg->lev[l].cell_start, kg->lev[l].bbox, d.N, TN,
                                                               c->c.reg_contact, c->contact, g_trans, c->lpart, c->LP, kg->resolved, l == 0);
                MH_LAUNCHED(c);
            }
            resolved = kg->resolved;
        }
#define MH_CONTACT(QQ) k_contact<QQ><<<mh_cdiv(TN, QQ), KNN_THREADS, 0, st>>>(c->verts, c->lowidx, c->scene, c->M, d.N, TN, c->c.reg_contact, c->contact, g_trans, c->lpart, c->LP, resolved)
        if (Q == 8) MH_CONTACT(8); else if (Q == 4) MH_CONTACT(4); else if (Q == 2) MH_CONTACT(2); else MH_CONTACT(1);
#undef MH_CONTACT
        MH_LAUNCHED(c);
        k_foot<<<mh_cdiv(d.T, d.B), 128, 0, st>>>(c->verts, c->lowidx, c->contact, d.T, d.N, d.B, c->c.reg_foot_sliding, c->dverts, c->lpart, c->LP);
        MH_LAUNCHED(c);
    }
    return MH_OK;
}

// -------------------------------------------------------------------------------------------------
// after the render stage: depth-range gradients and the depth / silhouette loss sums, one thread per frame
__global__ void k_frame_finalize(const float* __restrict__ pfout, const float* __restrict__ zmin_lin, const float* __restrict__ zmax_lin,
                                 int T, int N, float coef_depth, float* __restrict__ g_zmin, float* __restrict__ g_zmax,
                                 float* __restrict__ lpart, int LP) {
    __shared__ float sm[32];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float ld = 0.f, ls = 0.f;
    if (t < T) {
        float gmin = 0.f, gmax = 0.f;
        for (int n = 0; n < N; ++n) {
            const float* o = pfout + ((size_t)t * N + n) * PF_COUNT;
            const float inv = 1.0f / (o[PF_S] + 1.0f);
            const float a = o[PF_A] * inv, cc = o[PF_C] * inv;          // losses.py:24-27
            const float diff = a - cc;
            ld += diff * diff;
            ls += o[PF_SIL];
            const float k = -2.0f * diff * inv * coef_depth;              // dL/dc / (S + 1)
            gmin += k * o[PF_GIZMIN];
            gmax += k * o[PF_GIZMAX];
        }
        // min_z = log(1 + e^x) ; max_z = stopgrad(min_z) + 1 + log(1 + e^y)   (optimizer.py:683-688)
        const float ex = expf(zmin_lin[t]), ey = expf(zmax_lin[t]);
        const float minz = logf(1.0f + ex), maxz = minz + 1.0f + logf(1.0f + ey);
        g_zmin[t] += gmin * -(ex / (1.0f + ex)) / (minz * minz);
        g_zmax[t] += gmax * -(ey / (1.0f + ey)) / (maxz * maxz);
    }
    ld = block_sum(ld, sm);
    ls = block_sum(ls, sm);
    if (threadIdx.x == 0) { lpart[(size_t)MH_L_DEPTH * LP + blockIdx.x] = ld; lpart[(size_t)MH_L_SILHOUETTE * LP + blockIdx.x] = ls; }
}

// shape prior (weighted by the batch size per batch, :526) and the scale priors (once per batch, :531-542)
__global__ void k_shared_priors(const float* __restrict__ betas, const float* __restrict__ betas_ref, const float* __restrict__ xscale, int N,
                                int T_local, int nbatch_local, float coef_poses, float coef_scales, float* __restrict__ g_betas,
                                float* __restrict__ g_xscale, float* __restrict__ losses) {
    __shared__ float sm[32];
    const int tid = threadIdx.x;
    float lb = 0.f;
    for (int e = tid; e < N * MH_NBETA; e += blockDim.x) {
        const float df = betas[e] - betas_ref[e];
        lb += fabsf(df);
        g_betas[e] += coef_poses * (float)T_local * signf(df);
    }
    lb = block_sum(lb, sm);
    float s1 = 0.f, s2 = 0.f;
    for (int n = tid; n < N; n += blockDim.x) {
        const float s = powf(1.1f, xscale[n]) - 1.0f;
        s1 += s; s2 += s * s;
    }
    s1 = block_sum(s1, sm);
    __shared__ float ssum;
    if (tid == 0) ssum = s1;
    s2 = block_sum(s2, sm);
    __syncthreads();
    const float tot = ssum;
    for (int n = tid; n < N; n += blockDim.x) {
        const float sc = powf(1.1f, xscale[n]);
        const float g = coef_scales * 2.0f * (sc - 1.0f) / (float)N + (coef_scales > 0.f ? 1.0f : 0.0f) * 2.0f * tot;
        g_xscale[n] += (float)nbatch_local * g * 0.09531017980432493f * sc;
    }
    if (tid == 0) {                 // the only block, after k_loss_reduce in stream order: plain adds
        losses[MH_L_REF_POSES] += (float)T_local * lb;
        losses[MH_L_SCALE] += (float)nbatch_local * (tot * tot + s2 / (float)N);
    }
}

// losses[slot] += sum of the slot's partials, in a fixed order (strided partial sums per thread, then a fixed tree)
__global__ void __launch_bounds__(256) k_loss_reduce(const float* __restrict__ lpart, int LP, float* __restrict__ losses) {
    __shared__ float sm[256];
    const int slot = blockIdx.x, tid = threadIdx.x;
    float a = 0.f;
    for (int e = tid; e < LP; e += 256) a += lpart[(size_t)slot * LP + e];
    sm[tid] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) sm[tid] += sm[tid + o];
        __syncthreads();
    }
    if (tid == 0) losses[slot] += sm[0];
}

int mh_loss_begin(mh_ctx* c, cudaStream_t st) {
    MH_CUDA(c, cudaMemsetAsync(c->lpart, 0, sizeof(float) * (size_t)MH_L_COUNT * c->LP, st));
    return MH_OK;
}

int mh_loss_reduce(mh_ctx* c, cudaStream_t st) {
    k_loss_reduce<<<MH_L_COUNT, 256, 0, st>>>(c->lpart, c->LP, c->grads + c->n_params);
    MH_LAUNCHED(c);
    return MH_OK;
}

int mh_terms_post(mh_ctx* c, cudaStream_t st) {
    const mh_dims& d = c->d;
    float* losses = c->grads + c->n_params;
    if (c->c.depth != 0.f || c->c.silhouette != 0.f) {
        k_frame_finalize<<<mh_cdiv(d.T, 128), 128, 0, st>>>(c->pfout, c->params + c->off[MH_P_ZMIN_LIN], c->params + c->off[MH_P_ZMAX_LIN],
                                                            d.T, d.N, c->c.depth, c->grads + c->off[MH_P_ZMIN_LIN],
                                                            c->grads + c->off[MH_P_ZMAX_LIN], c->lpart, c->LP);
        MH_LAUNCHED(c);
    }
    MH_TRY(mh_loss_reduce(c, st));
    k_shared_priors<<<1, 128, 0, st>>>(c->params + c->off[MH_P_BETAS], c->betas_ref, c->params + c->off[MH_P_XSCALE], d.N, d.T,
                                       mh_cdiv(d.T, d.B), c->c.reg_poses, c->c.reg_scales, c->grads + c->off[MH_P_BETAS],
                                       c->grads + c->off[MH_P_XSCALE], losses);
    MH_LAUNCHED(c);
    return MH_OK;
}

// -------------------------------------------------------------------------------------------------
// hot loop A (optimizer.py:740-761): one warp per local person-frame.
//   loss_2d = mean over (T_total, N, 17, 2) of (vis * proj - vis * gt)^2   [pixels]
__global__ void k_init_2d(const float* __restrict__ j17, const float* __restrict__ vis, const float* __restrict__ pose2d,
                          const float* __restrict__ trans, const float* __restrict__ xscale, CamParams cam, int T, int N, float inv_count,
                          float coef_proj, float* __restrict__ g_trans, float* __restrict__ lpart, int LP) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= T * N) return;
    const float s = powf(1.1f, xscale[i % N]);
    float l = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (lane < MH_NJR) {
        const float* J = j17 + ((size_t)i * MH_NJR + lane) * 3;
        const float P[3] = {s * J[0] + trans[(size_t)i * 3], s * J[1] + trans[(size_t)i * 3 + 1], s * J[2] + trans[(size_t)i * 3 + 2]};
        float uv[2];
        mh_project(P, cam.K, cam.has_kd ? cam.Kd : nullptr, uv);
        const float v = vis[(size_t)i * MH_NJR + lane] * cam.w17[lane];            // pose_weights * vis_pose2d (optimizer.py:754-756)
        const float* q = pose2d + ((size_t)i * MH_NJR + lane) * 3;
        const float du = v * uv[0] - v * q[0], dv = v * uv[1] - v * q[1];
        l = du * du + dv * dv;
        float gP[3];
        mh_project_bwd(P, cam.K, cam.has_kd ? cam.Kd : nullptr, coef_proj * 2.0f * du * v * inv_count, coef_proj * 2.0f * dv * v * inv_count, gP);
        g0 = gP[0]; g1 = gP[1]; g2 = gP[2];
    }
    l = warp_sum(l); g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2);
    if (lane == 0) {
        lpart[(size_t)MH_L_INIT_2D * LP + i] = l;
        g_trans[(size_t)i * 3] += g0; g_trans[(size_t)i * 3 + 1] += g1; g_trans[(size_t)i * 3 + 2] += g2;
    }
}

int mh_init_iter_grads(mh_ctx* c, int use_prev, int use_next, cudaStream_t st) {
    const mh_dims& d = c->d;
    const int TN = d.T * d.N;
    if (d.t0 == 0) use_prev = 0;
    if (d.t0 + d.T == d.T_total) use_next = 0;
    float* losses = c->grads + c->n_params;
    float* g_trans = c->grads + c->off[MH_P_POSES_T];
    MH_CUDA(c, cudaMemsetAsync(g_trans, 0, sizeof(float) * TN * 3, st));
    MH_CUDA(c, cudaMemsetAsync(losses, 0, sizeof(float) * MH_L_COUNT, st));
    MH_TRY(mh_loss_begin(c, st));
    MH_TRY(mh_terms_gather(c, use_prev, use_next, st));
    const float inv_count = 1.0f / ((float)d.T_total * d.N * MH_NJR * 2);
    k_init_2d<<<mh_cdiv(TN, 4), 128, 0, st>>>(c->init_j17, c->init_vis, c->pose2d, c->params + c->off[MH_P_POSES_T],
                                              c->params + c->off[MH_P_XSCALE], cam_of(c), d.T, d.N, inv_count, c->c.proj2d, g_trans, c->lpart, c->LP);
    MH_LAUNCHED(c);
    k_velocity<<<mh_cdiv(TN * 3, 256), 256, 0, st>>>(c->trans_all, d.T, d.N, d.t0, d.T_total, use_prev, use_next, c->c.reg_velocity, g_trans,
                                                     c->lpart, c->LP);
    MH_LAUNCHED(c);
    MH_TRY(mh_loss_reduce(c, st));
    return MH_OK;
}
