// C ABI of libmhopt.so (include/mhopt.h): context lifetime, model / camera / frame ingest, parameter
// access, the fused optimiser updates and the per-cycle entry points.
//
// Reference code replaced (paths relative to the reference repo): SMPLOptimizerBase.__init__ and
// SMPLDepthSequenceOptimizer.__init__ / init_optimized_variables / fit / get_optimized_variables
// (mhmocap/optimizer.py:35-131, 150-321, 324-602, 619-636); torch.optim.RMSprop / Adam as instantiated at
// optimizer.py:355-356, 738-739.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "mh_ctx.h"

#define API_BEGIN(ctx) if (!(ctx)) return MH_E_ARG; cudaSetDevice((ctx)->d.device)

template <typename T>
static int dev_alloc(mh_ctx* c, T** p, int64_t n) {
    *p = nullptr;
    if (n <= 0) n = 1;
    cudaError_t e = mh_dev_alloc((void**)p, (size_t)n * sizeof(T));
    if (e != cudaSuccess) MH_FAIL(c, MH_E_CUDA, "mh_dev_alloc(%lld bytes): %s", (long long)(n * sizeof(T)), cudaGetErrorString(e));
    e = cudaMemset(*p, 0, (size_t)n * sizeof(T));
    if (e != cudaSuccess) MH_FAIL(c, MH_E_CUDA, "cudaMemset: %s", cudaGetErrorString(e));
    c->allocs.push_back((void*)*p);
    return MH_OK;
}

template <typename T>
static int upload(mh_ctx* c, T** p, const std::vector<T>& h) {
    MH_TRY(dev_alloc(c, p, (int64_t)h.size()));
    if (!h.empty()) MH_CUDA(c, cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MH_OK;
}

int mh_upload_floats(mh_ctx* c, float** p, const std::vector<float>& h) { return upload(c, p, h); }
int mh_alloc_floats(mh_ctx* c, float** p, int64_t n) { return dev_alloc(c, p, n); }      // zeroed, released by mh_destroy
int mh_alloc_ints(mh_ctx* c, int** p, int64_t n) { return dev_alloc(c, p, n); }

extern "C" const char* mh_version(void) { return "mhopt-b200 0.1 (sm_100a)"; }

extern "C" const char* mh_last_error(const mh_ctx* c) { return c ? c->err : "null context"; }

extern "C" int64_t mh_launch_count(const mh_ctx* c) { return c ? c->launches : 0; }

extern "C" int mh_set_batch(mh_ctx* c, int32_t B) {
    if (!c) return MH_E_ARG;
    const mh_dims& d = c->d;
    if (B < 1) MH_FAIL(c, MH_E_ARG, "mh_set_batch: batch size %d", B);
    // foot-sliding pairs never cross a batch (optimizer.py:512-518): shard edges must coincide with batch edges
    if (d.t0 % B != 0) MH_FAIL(c, MH_E_ARG, "shard start %d is not a multiple of the batch size %d", d.t0, B);
    if (d.t0 + d.T != d.T_total && d.T % B != 0) MH_FAIL(c, MH_E_ARG, "inner shard length %d is not a multiple of the batch size %d", d.T, B);
    c->d.B = B;
    return MH_OK;
}

extern "C" int mh_create(mh_ctx** out, const mh_dims* dims) {
    if (!out || !dims) return MH_E_ARG;
    *out = nullptr;
    mh_ctx* c = new mh_ctx();
    c->d = *dims;
    c->err[0] = 0;
    c->launches = 0;
    c->model_set = c->camera_set = c->coefs_set = c->ingested = c->has_filters = c->init_ready = false;
    c->optim_scale = true;
    c->rs = nullptr;
    c->scene_state = nullptr;
    c->scene_post = nullptr;
    c->knn = nullptr;
    c->scene_counts = nullptr;
    c->comm = nullptr;
    c->M = 0;
    c->events = nullptr; c->timing = false; c->timing_iter = 0;
    for (int k = 0; k < MH_NJR; ++k) c->w17[k] = 1.0f;
    *out = c;     // returned even on failure so that the caller can read the error text, then mh_destroy
    const mh_dims& d = c->d;
    if (d.T < 1 || d.N < 1 || d.N > MH_MAXN || d.H < 1 || d.W < 1 || d.V != MH_V || d.F != MH_F || d.B < 1 || d.world < 1 ||
        d.rank < 0 || d.rank >= d.world || d.t0 < 0 || d.t0 + d.T > d.T_total || d.M_max < 0)
        MH_FAIL(c, MH_E_ARG, "mh_create: bad dims (T=%d N=%d H=%d W=%d V=%d F=%d B=%d rank=%d/%d t0=%d T_total=%d)", d.T, d.N, d.H,
                d.W, d.V, d.F, d.B, d.rank, d.world, d.t0, d.T_total);
    MH_TRY(mh_set_batch(c, d.B));
    MH_CUDA(c, cudaSetDevice(d.device));
    cudaDeviceProp prop;
    MH_CUDA(c, cudaGetDeviceProperties(&prop, d.device));
    c->num_sms = prop.multiProcessorCount;
    c->Ts = d.T + 2;
    c->nb = c->Ts * d.N;
    const int64_t TN = (int64_t)d.T * d.N, HW = (int64_t)d.H * d.W, nb = c->nb;
    // parameter layout
    int64_t o = 0;
    c->off[MH_P_POSES_T] = o;    c->cnt[MH_P_POSES_T] = TN * 3;     o += TN * 3;
    c->off[MH_P_POSES_SMPL] = o; c->cnt[MH_P_POSES_SMPL] = TN * 72; o += TN * 72;
    c->off[MH_P_ZMIN_LIN] = o;   c->cnt[MH_P_ZMIN_LIN] = d.T;       o += d.T;
    c->off[MH_P_ZMAX_LIN] = o;   c->cnt[MH_P_ZMAX_LIN] = d.T;       o += d.T;
    c->off[MH_P_BETAS] = o;      c->cnt[MH_P_BETAS] = d.N * 10;     o += d.N * 10;
    c->off[MH_P_XSCALE] = o;     c->cnt[MH_P_XSCALE] = d.N;         o += d.N;
    c->off[MH_P_BETAS_REF] = -1; c->cnt[MH_P_BETAS_REF] = d.N * 10;
    c->n_params = o;
    MH_TRY(dev_alloc(c, &c->params, o));
    MH_TRY(dev_alloc(c, &c->grads, o + MH_L_COUNT));
    MH_TRY(dev_alloc(c, &c->sqavg, o));
    MH_TRY(dev_alloc(c, &c->mom, o));
    MH_TRY(dev_alloc(c, &c->betas_ref, d.N * 10));
    MH_TRY(dev_alloc(c, &c->halo_send, 2 * d.N * MH_HALO));
    MH_TRY(dev_alloc(c, &c->halo_recv, 2 * d.N * MH_HALO));
    MH_TRY(dev_alloc(c, &c->depth, d.T * HW));
    MH_TRY(dev_alloc(c, &c->cbits, d.T * HW));
    MH_TRY(dev_alloc(c, &c->ebits, d.T * HW));
    c->stage = nullptr; c->stage_floats = 0;
    c->pack_buf[0] = c->pack_buf[1] = nullptr; c->pack_ev[0] = c->pack_ev[1] = nullptr; c->pack_words = 0; c->pack_idx = 0; c->host_nonbinary = false;
    MH_TRY(dev_alloc(c, &c->pose2d, TN * 17 * 3));
    MH_TRY(dev_alloc(c, &c->theta_ref, TN * 72));
    MH_TRY(dev_alloc(c, &c->valid, TN));
    MH_TRY(dev_alloc(c, &c->maskarea, TN));
    MH_TRY(dev_alloc(c, &c->pose2d_valid, TN));
    MH_TRY(dev_alloc(c, &c->mask_valid, TN));
    MH_TRY(dev_alloc(c, &c->devflags, 8));
    MH_TRY(dev_alloc(c, &c->init_j17, TN * 17 * 3));
    MH_TRY(dev_alloc(c, &c->init_vis, TN * 17));
    MH_TRY(dev_alloc(c, &c->adam_m, TN * 3));
    MH_TRY(dev_alloc(c, &c->adam_v, TN * 3));
    MH_TRY(dev_alloc(c, &c->theta_all, nb * 72));
    MH_TRY(dev_alloc(c, &c->trans_all, nb * 3));
    MH_TRY(dev_alloc(c, &c->vshaped, (int64_t)d.N * MH_LD3V));
    MH_TRY(dev_alloc(c, &c->Jrest, nb * 72));
    MH_TRY(dev_alloc(c, &c->A, nb * 288));
    MH_TRY(dev_alloc(c, &c->pf, nb * MH_KPF));
    MH_TRY(dev_alloc(c, &c->vposed, nb * MH_LD3V));
    MH_TRY(dev_alloc(c, &c->verts, nb * MH_LD3V));
    MH_TRY(dev_alloc(c, &c->dverts, nb * MH_LD3V));
    MH_TRY(dev_alloc(c, &c->filtered, nb * MH_LD3V));
    MH_TRY(dev_alloc(c, &c->j17, nb * 17 * 3));
    MH_TRY(dev_alloc(c, &c->gj17, nb * 17 * 3));
    MH_TRY(dev_alloc(c, &c->lowidx, nb));
    MH_TRY(dev_alloc(c, &c->dA, nb * 288));
    MH_TRY(dev_alloc(c, &c->gT, nb * 4));
    MH_TRY(dev_alloc(c, &c->shared_part, TN * 12));
    c->LP = (int)(8 * TN);
    MH_TRY(dev_alloc(c, &c->lpart, (int64_t)MH_L_COUNT * c->LP));
    MH_TRY(dev_alloc(c, &c->dpf_part, (int64_t)(MH_KSPLIT + 1) * nb * MH_NEXT));
    MH_TRY(dev_alloc(c, &c->order, TN));
    MH_TRY(dev_alloc(c, &c->premask, TN));
    MH_TRY(dev_alloc(c, &c->dirty, d.T));
    MH_TRY(dev_alloc(c, &c->rankcnt, (int64_t)d.T * (d.N + 1)));
    MH_TRY(dev_alloc(c, &c->pfout, TN * PF_COUNT));
    MH_TRY(dev_alloc(c, &c->scene, std::max<int64_t>(d.M_max, 1) * 3));
    MH_TRY(dev_alloc(c, &c->contact, TN * 4));
    c->carry_floats = 2 * ((int64_t)d.N * MH_LD3V + d.N * 3);
    MH_TRY(dev_alloc(c, &c->carry_in, c->carry_floats));
    MH_TRY(dev_alloc(c, &c->carry_out, c->carry_floats));
    MH_TRY(dev_alloc(c, &c->transfilt, TN * 3));
    MH_TRY(dev_alloc(c, &c->pix_x, d.W));
    MH_TRY(dev_alloc(c, &c->pix_y, d.H));
    // order starts as "unknown" so that the first prepass recomputes every frame
    MH_CUDA(c, cudaMemset(c->order, 0xff, TN * sizeof(int)));
    MH_TRY(mh_render_alloc(c));
    MH_CUDA(c, cudaDeviceSynchronize());
    return MH_OK;
}

extern "C" void mh_destroy(mh_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->d.device);
    cudaDeviceSynchronize();
    mh_render_free(c);
    mh_scene_free(c);
    mh_scenepost_free(c);
    mh_knn_free(c);
    mh_comm_free(c);
    if (c->events) { for (int i = 0; i < MH_TIMING_RING * MH_TIMING_EVENTS; ++i) cudaEventDestroy(c->events[i]); delete[] c->events; }
    for (void* p : c->allocs) mh_dev_free(p);
    if (c->stage) mh_dev_free(c->stage);
    for (int i = 0; i < 2; ++i) { if (c->pack_buf[i]) cudaFreeHost(c->pack_buf[i]); if (c->pack_ev[i]) cudaEventDestroy(c->pack_ev[i]); }
    delete c;
}

// ---- model --------------------------------------------------------------------------------------
extern "C" int mh_set_model(mh_ctx* c, const mh_model* m) {
    API_BEGIN(c);
    if (!m || !m->v_template || !m->shapedirs || !m->posedirs || !m->J_regressor || !m->lbs_weights || !m->parents || !m->faces ||
        !m->reg17)
        MH_FAIL(c, MH_E_ARG, "mh_set_model: null buffer");
    static const int want[MH_NJ] = MH_PARENTS;
    for (int j = 0; j < MH_NJ; ++j)
        if (m->parents[j] != want[j]) MH_FAIL(c, MH_E_ARG, "mh_set_model: parents[%d] = %d, the SMPL tree has %d", j, m->parents[j], want[j]);
    const int V = MH_V;
    // extended basis: rows 0..188 live posedirs, 192..201 shapedirs (transposed to (l, 3v+k))
    std::vector<float> pext((size_t)MH_NEXT * MH_LD3V, 0.f);
    for (int r = 0; r < MH_NPF_LIVE; ++r) memcpy(&pext[(size_t)r * MH_LD3V], m->posedirs + (size_t)r * 3 * V, sizeof(float) * 3 * V);
    for (int v = 0; v < V; ++v)
        for (int k = 0; k < 3; ++k)
            for (int l = 0; l < MH_NBETA; ++l) pext[(size_t)(MH_KPF + l) * MH_LD3V + 3 * v + k] = m->shapedirs[((size_t)v * 3 + k) * MH_NBETA + l];
    std::vector<float> vt(MH_LD3V, 0.f);
    memcpy(vt.data(), m->v_template, sizeof(float) * 3 * V);
    // closed form of the rest joints: J = J_regressor.v_template + (J_regressor.shapedirs).beta  (smpl.py:532-535)
    std::vector<float> Jt(72, 0.f), Js(720, 0.f);
    for (int j = 0; j < MH_NJ; ++j) {
        double at[3] = {0, 0, 0};
        double as[3][MH_NBETA] = {{0}};
        for (int v = 0; v < V; ++v) {
            const float w = m->J_regressor[(size_t)j * V + v];
            if (w == 0.f) continue;
            for (int k = 0; k < 3; ++k) {
                at[k] += (double)w * m->v_template[3 * v + k];
                for (int l = 0; l < MH_NBETA; ++l) as[k][l] += (double)w * m->shapedirs[((size_t)v * 3 + k) * MH_NBETA + l];
            }
        }
        for (int k = 0; k < 3; ++k) {
            Jt[3 * j + k] = (float)at[k];
            for (int l = 0; l < MH_NBETA; ++l) Js[(3 * j + k) * MH_NBETA + l] = (float)as[k][l];
        }
    }
    // sparse skinning weights (exact: zero weights contribute nothing)
    int KW = 1;
    for (int v = 0; v < V; ++v) {
        int n = 0;
        for (int j = 0; j < MH_NJ; ++j) n += (m->lbs_weights[(size_t)v * MH_NJ + j] != 0.f);
        KW = std::max(KW, n);
    }
    std::vector<uint8_t> wj((size_t)V * KW, 0);
    std::vector<float> ww((size_t)V * KW, 0.f);
    std::vector<int> jptr(MH_NJ + 1, 0), jvert;
    std::vector<float> jw;
    for (int v = 0; v < V; ++v) {
        int q = 0;
        for (int j = 0; j < MH_NJ; ++j) {
            const float w = m->lbs_weights[(size_t)v * MH_NJ + j];
            if (w != 0.f) { wj[(size_t)v * KW + q] = (uint8_t)j; ww[(size_t)v * KW + q] = w; ++q; }
        }
    }
    for (int j = 0; j < MH_NJ; ++j) {
        for (int v = 0; v < V; ++v) {
            const float w = m->lbs_weights[(size_t)v * MH_NJ + j];
            if (w != 0.f) { jvert.push_back(v); jw.push_back(w); }
        }
        jptr[j + 1] = (int)jvert.size();
    }
    // 17-joint regressor: CSR (joint -> vertices) and CSC (vertex -> joints)
    std::vector<int> rptr(MH_NJR + 1, 0), rvert, cptr(V + 1, 0), cjoint;
    std::vector<float> rw, cw;
    for (int k = 0; k < MH_NJR; ++k) {
        for (int v = 0; v < V; ++v) {
            const float w = m->reg17[(size_t)k * V + v];
            if (w != 0.f) { rvert.push_back(v); rw.push_back(w); }
        }
        rptr[k + 1] = (int)rvert.size();
        double rs = 0.0;
        for (int e = rptr[k]; e < rptr[k + 1]; ++e) rs += (double)rw[e];
        c->r17_slack[k] = (float)(1.0 - rs);
    }
    for (int v = 0; v < V; ++v) {
        for (int k = 0; k < MH_NJR; ++k) {
            const float w = m->reg17[(size_t)k * V + v];
            if (w != 0.f) { cjoint.push_back(k); cw.push_back(w); }
        }
        cptr[v + 1] = (int)cjoint.size();
    }
    for (int f = 0; f < MH_F * 3; ++f)
        if (m->faces[f] < 0 || m->faces[f] >= V) MH_FAIL(c, MH_E_ARG, "mh_set_model: face index %d out of range", m->faces[f]);
    std::vector<int32_t> faces(m->faces, m->faces + (size_t)MH_F * 3);
    c->KW = KW; c->jnnz = (int)jvert.size(); c->rnnz = (int)rvert.size();
    MH_TRY(upload(c, &c->pext, pext));
    MH_TRY(mh_gemm_tc_prepare(c));
    MH_TRY(upload(c, &c->vtemplate, vt));
    MH_TRY(upload(c, &c->Jt, Jt));
    MH_TRY(upload(c, &c->Js, Js));
    MH_TRY(upload(c, &c->wj, wj));
    MH_TRY(upload(c, &c->ww, ww));
    MH_TRY(upload(c, &c->jptr, jptr));
    MH_TRY(upload(c, &c->jvert, jvert));
    MH_TRY(upload(c, &c->jw, jw));
    MH_TRY(upload(c, &c->rptr, rptr));
    MH_TRY(upload(c, &c->rvert, rvert));
    MH_TRY(upload(c, &c->rw, rw));
    MH_TRY(upload(c, &c->cptr, cptr));
    MH_TRY(upload(c, &c->cjoint, cjoint));
    MH_TRY(upload(c, &c->cw, cw));
    MH_TRY(upload(c, &c->faces, faces));
    c->model_set = true;
    return MH_OK;
}

// NDC coordinate of pixel-centre index i on an axis of S1 pixels (S2 = the other axis): PyTorch3D's
// non-square convention, evaluated in fp32 exactly like oracle.raster.pixel_centers_ndc.
static float ndc_center(int i, int S1, int S2) {
    const float r = (S1 > S2) ? (float)(2.0 * S1 / S2) : 2.0f;
    const float half = r / 2.0f;
    const float a = r * (float)i;
    const float b = a + half;
    const float q = b / (float)S1;
    return -half + q;
}

extern "C" int mh_set_camera(mh_ctx* c, const float K[9], const float Kndc[16], const float* Kd) {
    API_BEGIN(c);
    if (!K || !Kndc) MH_FAIL(c, MH_E_ARG, "mh_set_camera: null matrix");
    memcpy(c->K, K, sizeof(float) * 9);
    memcpy(c->Kndc, Kndc, sizeof(float) * 16);
    c->has_kd = Kd != nullptr;
    if (Kd) memcpy(c->Kd, Kd, sizeof(float) * 5);
    const int W = c->d.W, H = c->d.H;
    std::vector<float> px(W), py(H);
    for (int x = 0; x < W; ++x) px[x] = ndc_center(W - 1 - x, W, H);
    for (int y = 0; y < H; ++y) py[y] = ndc_center(H - 1 - y, H, W);
    MH_CUDA(c, cudaMemcpy(c->pix_x, px.data(), sizeof(float) * W, cudaMemcpyHostToDevice));
    MH_CUDA(c, cudaMemcpy(c->pix_y, py.data(), sizeof(float) * H, cudaMemcpyHostToDevice));
    c->camera_set = true;
    return MH_OK;
}

extern "C" int mh_set_joint_weights(mh_ctx* c, const float* w) {
    API_BEGIN(c);
    if (!w) MH_FAIL(c, MH_E_ARG, "mh_set_joint_weights: null");
    memcpy(c->w17, w, sizeof(c->w17));
    return MH_OK;
}

extern "C" int mh_set_coefs(mh_ctx* c, const mh_coefs* k) {
    API_BEGIN(c);
    if (!k) MH_FAIL(c, MH_E_ARG, "mh_set_coefs: null");
    c->c = *k;
    c->coefs_set = true;
    return MH_OK;
}

// ---- frames -------------------------------------------------------------------------------------
static int ingest_frames(mh_ctx* c, int32_t t0, int32_t count, const float* depths, const void* seg, int seg_is_u8, const float* pose2d,
                         const float* theta_ref, const float* valid, void* stream);

extern "C" int mh_ingest_frames(mh_ctx* c, int32_t t0, int32_t count, const float* depths, const float* seg, const float* pose2d,
                                const float* theta_ref, const float* valid, void* stream) {
    return ingest_frames(c, t0, count, depths, seg, 0, pose2d, theta_ref, valid, stream);
}

extern "C" int mh_ingest_frames_u8(mh_ctx* c, int32_t t0, int32_t count, const float* depths, const uint8_t* seg, const float* pose2d,
                                   const float* theta_ref, const float* valid, void* stream) {
    return ingest_frames(c, t0, count, depths, seg, 1, pose2d, theta_ref, valid, stream);
}

// float32 {0., 1.} masks (count, N, HW) -> one 32-bit plane per frame on the HOST, all cores: the reference dataset delivers 4 N bytes
// per pixel and frame (utils.py:329-331); packed, 4 bytes cross the bus instead.  Returns true when a value other than 0 / 1 was seen.
// one person's row of a chunk: the widest vector unit of the host is picked at load time (function multi-versioning)
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
static int mh_pack_row(const float* s, uint32_t* o, int64_t n, uint32_t bit) {
    int bad = 0;
    for (int64_t p = 0; p < n; ++p) {
        const float v = s[p];
        o[p] |= (v != 0.f) ? bit : 0u;
        bad |= (v != 0.f) & (v != 1.0f);
    }
    return bad;
}

static bool pack_masks_host(const float* seg, int count, int N, int64_t HW, uint32_t* out) {
    const int64_t total = (int64_t)count * HW;
    const int nthr = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    const int64_t chunk = 2048;                                   // pixels per work item: the 8 KB output chunk stays in L1 over the N passes
    std::atomic<int64_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        int mybad = 0;
        for (;;) {
            const int64_t i0 = next.fetch_add(chunk);
            if (i0 >= total) break;
            const int64_t i1 = std::min(total, i0 + chunk);
            int64_t i = i0;
            while (i < i1) {                                      // a chunk may straddle a frame boundary
                const int64_t t = i / HW, p0 = i - t * HW, n_here = std::min(i1 - i, HW - p0);
                uint32_t* o = out + i;
                for (int64_t p = 0; p < n_here; ++p) o[p] = 0u;
                for (int n = 0; n < N; ++n) {
                    mybad |= mh_pack_row(seg + ((int64_t)t * N + n) * HW + p0, o, n_here, 1u << n);
                }
                i += n_here;
            }
        }
        if (mybad) bad.store(1);
    };
    std::vector<std::thread> th;
    for (int k = 1; k < nthr; ++k) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return bad.load() != 0;
}

// testing aid (no device involved): the host packing of float32 masks (count, N, HW) into one 32-bit plane per frame; returns 1 when
// a value other than 0 / 1 was seen, 0 otherwise, MH_E_ARG on bad arguments
extern "C" int mh_debug_pack_masks(const float* seg, int32_t count, int32_t N, int64_t HW, uint32_t* out) {
    if (!seg || !out || count < 1 || N < 1 || N > MH_MAXN || HW < 1) return MH_E_ARG;
    return pack_masks_host(seg, count, N, HW, out) ? 1 : 0;
}

static int ingest_frames(mh_ctx* c, int32_t t0, int32_t count, const float* depths, const void* seg, int seg_is_u8, const float* pose2d,
                         const float* theta_ref, const float* valid, void* stream) {
    API_BEGIN(c);
    cudaStream_t st = (cudaStream_t)stream;
    const mh_dims& d = c->d;
    if (t0 < 0 || count < 1 || t0 + count > d.T) MH_FAIL(c, MH_E_ARG, "mh_ingest_frames: frames [%d, %d) outside [0, %d)", t0, t0 + count, d.T);
    if (!pose2d || !theta_ref || !valid) MH_FAIL(c, MH_E_ARG, "mh_ingest_frames: null buffer");
    const int64_t HW = (int64_t)d.H * d.W, nseg = (int64_t)count * d.N * HW, need = seg_is_u8 ? (nseg + 3) / 4 : nseg;
    // float32 masks are compacted on the host (MH_INGEST_HOST_PACK=0: on the device after a full-size copy, for A/B measurements)
    // -- only when this process is alone on the host: with one rank per GPU the ranks' copies run in parallel over their own PCIe
    // links while the packing threads would all share the same cores and memory channels
    const char* hp_env = getenv("MH_INGEST_HOST_PACK");
    const int host_pack_env = hp_env ? atoi(hp_env) : -1;
    // (measured on C3 with pinned masks: host packing 0.20 s, full-size copies + device packing 0.32 s = 53 GB/s over the bus, and a
    // batch-by-batch mix of the two -- unpacked whenever the copy engine is idle -- 0.22 s: both paths read the same 17 GB of host
    // memory, which is the limit on this host; the mix was dropped)
    const bool host_pack = host_pack_env >= 0 ? host_pack_env != 0 : d.world == 1;
    if (seg && !seg_is_u8 && host_pack) {
        const int64_t words = (int64_t)count * HW;
        if (words > c->pack_words) {
            MH_CUDA(c, cudaStreamSynchronize(st));
            for (int i = 0; i < 2; ++i) {
                if (c->pack_buf[i]) cudaFreeHost(c->pack_buf[i]);
                c->pack_buf[i] = nullptr;
                MH_CUDA(c, cudaHostAlloc((void**)&c->pack_buf[i], sizeof(uint32_t) * words, cudaHostAllocDefault));
                if (!c->pack_ev[i]) MH_CUDA(c, cudaEventCreateWithFlags(&c->pack_ev[i], cudaEventDisableTiming));
            }
            c->pack_words = words;
        }
        const int bi = c->pack_idx;
        c->pack_idx ^= 1;
        MH_CUDA(c, cudaEventSynchronize(c->pack_ev[bi]));           // the copy that last read this buffer is done (no-op the first time)
        if (depths) MH_CUDA(c, cudaMemcpyAsync(c->depth + (int64_t)t0 * HW, depths, sizeof(float) * count * HW, cudaMemcpyHostToDevice, st));
        if (pack_masks_host(reinterpret_cast<const float*>(seg), count, d.N, HW, c->pack_buf[bi])) c->host_nonbinary = true;
        MH_CUDA(c, cudaMemcpyAsync(c->cbits + (int64_t)t0 * HW, c->pack_buf[bi], sizeof(uint32_t) * words, cudaMemcpyHostToDevice, st));
        MH_CUDA(c, cudaEventRecord(c->pack_ev[bi], st));
        MH_CUDA(c, cudaMemcpyAsync(c->pose2d + (int64_t)t0 * d.N * 51, pose2d, sizeof(float) * count * d.N * 51, cudaMemcpyHostToDevice, st));
        MH_CUDA(c, cudaMemcpyAsync(c->theta_ref + (int64_t)t0 * d.N * 72, theta_ref, sizeof(float) * count * d.N * 72, cudaMemcpyHostToDevice, st));
        MH_CUDA(c, cudaMemcpyAsync(c->valid + (int64_t)t0 * d.N, valid, sizeof(float) * count * d.N, cudaMemcpyHostToDevice, st));
        return MH_OK;
    }
    if (seg && need > c->stage_floats) {
        MH_CUDA(c, cudaStreamSynchronize(st));
        if (c->stage) mh_dev_free(c->stage);
        c->stage = nullptr; c->stage_floats = 0;
        MH_CUDA(c, mh_dev_alloc((void**)&c->stage, need * sizeof(float)));
        c->stage_floats = need;
    }
    if (depths) MH_CUDA(c, cudaMemcpyAsync(c->depth + (int64_t)t0 * HW, depths, sizeof(float) * count * HW, cudaMemcpyHostToDevice, st));
    if (seg) MH_CUDA(c, cudaMemcpyAsync(c->stage, seg, seg_is_u8 ? (size_t)nseg : sizeof(float) * (size_t)nseg, cudaMemcpyHostToDevice, st));
    MH_CUDA(c, cudaMemcpyAsync(c->pose2d + (int64_t)t0 * d.N * 51, pose2d, sizeof(float) * count * d.N * 51, cudaMemcpyHostToDevice, st));
    MH_CUDA(c, cudaMemcpyAsync(c->theta_ref + (int64_t)t0 * d.N * 72, theta_ref, sizeof(float) * count * d.N * 72, cudaMemcpyHostToDevice, st));
    MH_CUDA(c, cudaMemcpyAsync(c->valid + (int64_t)t0 * d.N, valid, sizeof(float) * count * d.N, cudaMemcpyHostToDevice, st));
    if (seg) MH_TRY(mh_ingest_compact(c, t0, count, seg_is_u8, st));
    return MH_OK;
}

extern "C" int mh_finalize_ingest(mh_ctx* c, void* stream) {
    API_BEGIN(c);
    if (!c->coefs_set) MH_FAIL(c, MH_E_STATE, "mh_finalize_ingest: mh_set_coefs first (the joint confidence threshold is needed)");
    cudaStream_t st = (cudaStream_t)stream;
    MH_TRY(mh_ingest_derive(c, st));
    int flags[8];
    MH_CUDA(c, cudaMemcpyAsync(flags, c->devflags, sizeof(flags), cudaMemcpyDeviceToHost, st));
    MH_CUDA(c, cudaStreamSynchronize(st));
    if (flags[0] || c->host_nonbinary) {
        c->host_nonbinary = false;
        cudaMemsetAsync(c->devflags, 0, sizeof(int), st);
        MH_FAIL(c, MH_E_ARG, "ingest: seg_mask holds values other than 0 / 1 (instance masks must be binary, utils.py:329-331)");
    }
    c->ingested = true;
    return MH_OK;
}

extern "C" int mh_set_scene(mh_ctx* c, const float* pcd, int64_t M, void* stream) {
    API_BEGIN(c);
    if (M < 0 || M > c->d.M_max) MH_FAIL(c, MH_E_CAPACITY, "mh_set_scene: %lld points exceed M_max = %lld", (long long)M, (long long)c->d.M_max);
    if (M > 0 && M < MH_KNN) MH_FAIL(c, MH_E_ARG, "mh_set_scene: the contact term needs at least %d scene points", MH_KNN);
    if (M > 0) {
        if (!pcd) MH_FAIL(c, MH_E_ARG, "mh_set_scene: null cloud");
        MH_CUDA(c, cudaMemcpyAsync(c->scene, pcd, sizeof(float) * 3 * M, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    }
    c->M = M;
    return mh_knn_build(c, (cudaStream_t)stream);
}

extern "C" int mh_set_scene_from_depth(mh_ctx* c, const float* depth_host, const uint8_t* mask_host, void* stream) {
    API_BEGIN(c);
    if (!c->camera_set) MH_FAIL(c, MH_E_STATE, "mh_set_scene_from_depth: mh_set_camera first");
    if (!depth_host || !mask_host) MH_FAIL(c, MH_E_ARG, "mh_set_scene_from_depth: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)c->d.H * c->d.W;
    float* dd; uint8_t* dm;
    MH_CUDA(c, mh_dev_alloc((void**)&dd, HW * sizeof(float)));
    MH_CUDA(c, mh_dev_alloc((void**)&dm, HW));
    int r = MH_OK;
    if (cudaMemcpyAsync(dd, depth_host, HW * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(dm, mask_host, HW, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "mh_set_scene_from_depth: copy failed");
        r = MH_E_CUDA;
    }
    if (r == MH_OK) r = mh_scene_from_depth(c, dd, dm, st);
    cudaStreamSynchronize(st);
    mh_dev_free(dd); mh_dev_free(dm);
    return r;
}

// ---- parameters ---------------------------------------------------------------------------------
static int param_ptr(mh_ctx* c, int which, int64_t count, float** p, bool grad) {
    if (which < 0 || which >= MH_P_COUNT) MH_FAIL(c, MH_E_ARG, "unknown parameter id %d", which);
    if (count != c->cnt[which]) MH_FAIL(c, MH_E_ARG, "parameter %d holds %lld floats, caller passed %lld", which, (long long)c->cnt[which], (long long)count);
    if (which == MH_P_BETAS_REF) {
        if (grad) MH_FAIL(c, MH_E_ARG, "betas_ref has no gradient");
        *p = c->betas_ref;
    } else {
        *p = (grad ? c->grads : c->params) + c->off[which];
    }
    return MH_OK;
}

extern "C" int mh_set_param(mh_ctx* c, int which, const float* src, int64_t count, void* stream) {
    API_BEGIN(c);
    float* p;
    MH_TRY(param_ptr(c, which, count, &p, false));
    if (!src) MH_FAIL(c, MH_E_ARG, "mh_set_param: null source");
    MH_CUDA(c, cudaMemcpyAsync(p, src, sizeof(float) * count, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return MH_OK;
}

extern "C" int mh_get_param(mh_ctx* c, int which, float* dst, int64_t count) {
    API_BEGIN(c);
    float* p;
    MH_TRY(param_ptr(c, which, count, &p, false));
    MH_CUDA(c, cudaDeviceSynchronize());
    MH_CUDA(c, cudaMemcpy(dst, p, sizeof(float) * count, cudaMemcpyDeviceToHost));
    return MH_OK;
}

extern "C" int mh_get_grad(mh_ctx* c, int which, float* dst, int64_t count) {
    API_BEGIN(c);
    float* p;
    MH_TRY(param_ptr(c, which, count, &p, true));
    MH_CUDA(c, cudaDeviceSynchronize());
    MH_CUDA(c, cudaMemcpy(dst, p, sizeof(float) * count, cudaMemcpyDeviceToHost));
    return MH_OK;
}

extern "C" int mh_set_optimize_scale(mh_ctx* c, int32_t on) {
    API_BEGIN(c);
    c->optim_scale = on != 0;
    return MH_OK;
}

extern "C" int mh_device_view(mh_ctx* c, int which, void** ptr, int64_t* n) {
    API_BEGIN(c);
    if (!ptr || !n) MH_FAIL(c, MH_E_ARG, "mh_device_view: null output");
    const mh_dims& d = c->d;
    switch (which) {
        case MH_BUF_SHARED: *ptr = c->grads + c->off[MH_P_BETAS]; *n = d.N * 11 + MH_L_COUNT; break;
        case MH_BUF_HALO_SEND: *ptr = c->halo_send; *n = 2 * d.N * MH_HALO; break;
        case MH_BUF_HALO_RECV: *ptr = c->halo_recv; *n = 2 * d.N * MH_HALO; break;
        case MH_BUF_CARRY_OUT: *ptr = c->carry_out; *n = c->carry_floats; break;
        case MH_BUF_CARRY_IN: *ptr = c->carry_in; *n = c->carry_floats; break;
        case MH_BUF_GRADS: *ptr = c->grads; *n = c->n_params + MH_L_COUNT; break;
        case MH_BUF_VERTS: *ptr = c->verts; *n = (int64_t)c->nb * MH_LD3V; break;
        case MH_BUF_FILTERED: *ptr = c->filtered; *n = (int64_t)c->nb * MH_LD3V; break;
        case MH_BUF_PARAMS: *ptr = c->params; *n = c->n_params; break;
        case MH_BUF_MEDIAN_HIST: case MH_BUF_MEDIAN_AUX: return mh_scene_views(c, which, ptr, n);
        default: MH_FAIL(c, MH_E_ARG, "mh_device_view: unknown buffer %d", which);
    }
    return MH_OK;
}

// ---- fused optimiser updates -----------------------------------------------------------------------
// torch.optim.RMSprop(lr, alpha=.5, eps=1e-8, momentum=.9, centered=False) (optimizer.py:355):
//   v <- alpha v + (1 - alpha) g^2 ; buf <- mu buf + g / (sqrt(v) + eps) ; p <- p - lr buf
__global__ void k_rmsprop(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, float* __restrict__ buf,
                          int64_t n, int64_t skip_lo, int64_t skip_hi, float lr, float alpha, float mu, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (i >= skip_lo && i < skip_hi)) return;
    const float gi = g[i];
    const float vi = v[i] * alpha + (1.0f - alpha) * gi * gi;
    const float bi = buf[i] * mu + gi / (sqrtf(vi) + eps);
    v[i] = vi; buf[i] = bi;
    p[i] = p[i] - lr * bi;
}

// torch.optim.Adam(lr, betas=(b1, b2), eps) with bias correction (optimizer.py:738)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                       float step_size, float b1, float b2, float inv_sqrt_bc2, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = m[i] + (1.0f - b1) * (gi - m[i]);          // lerp_
    const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] - step_size * (mi / denom);
}

extern "C" int mh_reset_optimizer(mh_ctx* c, void* stream) {
    API_BEGIN(c);
    cudaStream_t st = (cudaStream_t)stream;
    MH_CUDA(c, cudaMemsetAsync(c->sqavg, 0, sizeof(float) * c->n_params, st));
    MH_CUDA(c, cudaMemsetAsync(c->mom, 0, sizeof(float) * c->n_params, st));
    return MH_OK;
}

extern "C" int mh_fit_update(mh_ctx* c, float lr, void* stream) {
    API_BEGIN(c);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t lo = 0, hi = 0;
    if (!c->optim_scale) { lo = c->off[MH_P_XSCALE]; hi = lo + c->d.N; }      // "Not optimizing scale_factor" (optimizer.py:350-353)
    k_rmsprop<<<mh_cdiv(c->n_params, 256), 256, 0, st>>>(c->params, c->grads, c->sqavg, c->mom, c->n_params, lo, hi, lr, 0.5f, 0.9f, 1e-8f);
    MH_LAUNCHED(c);
    return MH_OK;
}

extern "C" int mh_init_update(mh_ctx* c, float lr, int32_t step, void* stream) {
    API_BEGIN(c);
    if (step < 1) MH_FAIL(c, MH_E_ARG, "mh_init_update: step is 1-based");
    cudaStream_t st = (cudaStream_t)stream;
    const double b1 = 0.5, b2 = 0.5;
    const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    const int64_t n = c->cnt[MH_P_POSES_T];
    k_adam<<<mh_cdiv(n, 256), 256, 0, st>>>(c->params + c->off[MH_P_POSES_T], c->grads + c->off[MH_P_POSES_T], c->adam_m, c->adam_v, n,
                                            (float)(lr / bc1), (float)b1, (float)b2, (float)(1.0 / sqrt(bc2)), 1e-6f);
    MH_LAUNCHED(c);
    return MH_OK;
}

// ---- SMPL forward utility -----------------------------------------------------------------------
extern "C" int mh_smpl_forward(mh_ctx* c, const float* betas, const float* theta, int64_t nbodies, float* verts, float* joints17) {
    API_BEGIN(c);
    if (!c->model_set) MH_FAIL(c, MH_E_STATE, "mh_smpl_forward: mh_set_model first");
    if (!betas || !theta || nbodies < 1) MH_FAIL(c, MH_E_ARG, "mh_smpl_forward: bad arguments");
    cudaStream_t st = 0;
    // scratch: the per-iteration arrays of the context (overwritten by the next cycle anyway); per-body shapes live in `dverts`
    const int64_t chunk = c->nb;
    float* dbetas;
    MH_CUDA(c, mh_dev_alloc((void**)&dbetas, sizeof(float) * chunk * 10));
    std::vector<float> hv;
    int r = MH_OK;
    for (int64_t s = 0; s < nbodies && r == MH_OK; s += chunk) {
        const int n = (int)std::min<int64_t>(chunk, nbodies - s);
        cudaMemcpyAsync(dbetas, betas + s * 10, sizeof(float) * n * 10, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(c->theta_all, theta + s * 72, sizeof(float) * n * 72, cudaMemcpyHostToDevice, st);
        MhSmplArgs a = {dbetas, n, 1, c->theta_all, nullptr, nullptr, n, c->d.N, c->dverts, c->Jrest, c->A, c->pf, c->vposed,
                        c->verts, c->j17, nullptr};
        r = mh_smpl_forward_run(c, a, st);
        if (r != MH_OK) break;
        if (verts) {
            if (cudaMemcpy2DAsync(verts + s * 3 * MH_V, sizeof(float) * 3 * MH_V, c->verts, sizeof(float) * MH_LD3V, sizeof(float) * 3 * MH_V, n,
                                  cudaMemcpyDeviceToHost, st) != cudaSuccess) r = MH_E_CUDA;
        }
        if (joints17 && cudaMemcpyAsync(joints17 + s * 51, c->j17, sizeof(float) * n * 51, cudaMemcpyDeviceToHost, st) != cudaSuccess) r = MH_E_CUDA;
        if (cudaStreamSynchronize(st) != cudaSuccess) r = MH_E_CUDA;
    }
    mh_dev_free(dbetas);
    if (r == MH_E_CUDA && !c->err[0]) snprintf(c->err, sizeof(c->err), "mh_smpl_forward: %s", cudaGetErrorString(cudaGetLastError()));
    return r;
}

// ---- evaluation: SMPL forward + an arbitrary sparse joint regressor ---------------------------------
// joints[b][j] = sum_v R[j][v] verts_local[b][v]   (smpl.py:376-389: vertices2joints with J_regressor_mupots / _h36m17 / _extra9;
// evaluate.py:222-229).  One warp per (body, joint) walks the joint's non-zeros.
__global__ void k_regress_joints(const float* __restrict__ verts, const int* __restrict__ rptr, const int* __restrict__ rvert,
                                 const float* __restrict__ rw, int nbodies, int J, float* __restrict__ out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nbodies * J) return;
    const int b = w / J, j = w % J;
    const float* v = verts + (size_t)b * MH_LD3V;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int e = rptr[j] + lane; e < rptr[j + 1]; e += 32) {
        const float wt = rw[e];
        const int vi = rvert[e];
        ax = fmaf(wt, v[3 * vi], ax); ay = fmaf(wt, v[3 * vi + 1], ay); az = fmaf(wt, v[3 * vi + 2], az);
    }
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_down_sync(0xffffffffu, ax, o); ay += __shfl_down_sync(0xffffffffu, ay, o); az += __shfl_down_sync(0xffffffffu, az, o);
    }
    if (lane == 0) { float* o = out + ((size_t)b * J + j) * 3; o[0] = ax; o[1] = ay; o[2] = az; }
}

extern "C" int mh_smpl_regress(mh_ctx* c, const float* betas, const float* theta, int64_t nbodies, const float* regressor, int32_t J,
                               float* joints) {
    API_BEGIN(c);
    if (!c->model_set) MH_FAIL(c, MH_E_STATE, "mh_smpl_regress: mh_set_model first");
    if (!betas || !theta || !regressor || !joints || nbodies < 1 || J < 1 || J > 64) MH_FAIL(c, MH_E_ARG, "mh_smpl_regress: bad arguments");
    cudaStream_t st = 0;
    // CSR of the (J, V) regressor: zero weights dropped (exact)
    std::vector<int> rptr(J + 1, 0), rvert;
    std::vector<float> rw;
    for (int j = 0; j < J; ++j) {
        for (int v = 0; v < MH_V; ++v) {
            const float w = regressor[(size_t)j * MH_V + v];
            if (w != 0.f) { rvert.push_back(v); rw.push_back(w); }
        }
        rptr[j + 1] = (int)rvert.size();
    }
    if (rvert.empty()) { rvert.push_back(0); rw.push_back(0.f); }
    const int64_t chunk = c->nb;
    int *d_rptr = nullptr, *d_rvert = nullptr;
    float *d_rw = nullptr, *dbetas = nullptr, *dj = nullptr;
    int r = MH_OK;
    if (mh_dev_alloc((void**)&d_rptr, sizeof(int) * rptr.size()) != cudaSuccess || mh_dev_alloc((void**)&d_rvert, sizeof(int) * rvert.size()) != cudaSuccess ||
        mh_dev_alloc((void**)&d_rw, sizeof(float) * rw.size()) != cudaSuccess || mh_dev_alloc((void**)&dbetas, sizeof(float) * chunk * 10) != cudaSuccess ||
        mh_dev_alloc((void**)&dj, sizeof(float) * chunk * J * 3) != cudaSuccess)
        r = MH_E_CUDA;
    if (r == MH_OK) {
        cudaMemcpyAsync(d_rptr, rptr.data(), sizeof(int) * rptr.size(), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_rvert, rvert.data(), sizeof(int) * rvert.size(), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_rw, rw.data(), sizeof(float) * rw.size(), cudaMemcpyHostToDevice, st);
    }
    for (int64_t s = 0; s < nbodies && r == MH_OK; s += chunk) {
        const int n = (int)std::min<int64_t>(chunk, nbodies - s);
        cudaMemcpyAsync(dbetas, betas + s * 10, sizeof(float) * n * 10, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(c->theta_all, theta + s * 72, sizeof(float) * n * 72, cudaMemcpyHostToDevice, st);
        // scale 1, translation 0 (null pointers): local vertices, per-body shapes in `dverts` as in mh_smpl_forward
        MhSmplArgs a = {dbetas, n, 1, c->theta_all, nullptr, nullptr, n, c->d.N, c->dverts, c->Jrest, c->A, c->pf, c->vposed,
                        c->verts, c->j17, nullptr};
        r = mh_smpl_forward_run(c, a, st);
        if (r != MH_OK) break;
        k_regress_joints<<<mh_cdiv((int64_t)n * J * 32, 256), 256, 0, st>>>(c->verts, d_rptr, d_rvert, d_rw, n, J, dj);
        c->launches++;
        if (cudaMemcpyAsync(joints + s * J * 3, dj, sizeof(float) * n * J * 3, cudaMemcpyDeviceToHost, st) != cudaSuccess) r = MH_E_CUDA;
        if (cudaStreamSynchronize(st) != cudaSuccess) r = MH_E_CUDA;
    }
    mh_dev_free(d_rptr); mh_dev_free(d_rvert); mh_dev_free(d_rw); mh_dev_free(dbetas); mh_dev_free(dj);
    if (r == MH_E_CUDA && !c->err[0]) snprintf(c->err, sizeof(c->err), "mh_smpl_regress: %s", cudaGetErrorString(cudaGetLastError()));
    return r;
}

// ---- hot loop A ---------------------------------------------------------------------------------
__global__ void k_init_fill(float* __restrict__ poses_T, int64_t n, const float* __restrict__ pose2d, float* __restrict__ vis, int64_t nj,
                            float thr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) poses_T[i] = (i % 3 == 2) ? 1.0f : 0.0f;                              // optimizer.py:729
    if (i < nj) vis[i] = pose2d[i * 3 + 2] > thr ? 1.0f : 0.0f;                       // optimizer.py:735
}

extern "C" int mh_init_begin(mh_ctx* c, const float* pose2d, const float* theta, const float* betas, float joints_thr, void* stream) {
    API_BEGIN(c);
    if (!c->model_set || !c->camera_set || !c->coefs_set) MH_FAIL(c, MH_E_STATE, "mh_init_begin: set model, camera and coefficients first");
    if (!pose2d || !theta || !betas) MH_FAIL(c, MH_E_ARG, "mh_init_begin: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const mh_dims& d = c->d;
    const int TN = d.T * d.N;
    float* dbetas;
    MH_CUDA(c, mh_dev_alloc((void**)&dbetas, sizeof(float) * TN * 10));
    MH_CUDA(c, cudaMemcpyAsync(dbetas, betas, sizeof(float) * TN * 10, cudaMemcpyHostToDevice, st));
    MH_CUDA(c, cudaMemcpyAsync(c->theta_all, theta, sizeof(float) * TN * 72, cudaMemcpyHostToDevice, st));
    MH_CUDA(c, cudaMemcpyAsync(c->pose2d, pose2d, sizeof(float) * TN * 51, cudaMemcpyHostToDevice, st));
    // the regressed joints do not depend on the optimised translation: evaluate SMPL once (the reference
    // re-evaluates it every iteration, optimizer.py:746-748) with the PER-FRAME betas (optimizer.py:733)
    MhSmplArgs a = {dbetas, TN, 1, c->theta_all, nullptr, nullptr, TN, d.N, c->dverts, c->Jrest, c->A, c->pf, c->vposed, c->verts,
                    c->init_j17, nullptr};
    int r = mh_smpl_forward_run(c, a, st);
    if (r == MH_OK) {
        k_init_fill<<<mh_cdiv((int64_t)TN * 17, 256), 256, 0, st>>>(c->params + c->off[MH_P_POSES_T], (int64_t)TN * 3, c->pose2d, c->init_vis,
                                                                  (int64_t)TN * 17, joints_thr);
        c->launches++;
        cudaMemsetAsync(c->adam_m, 0, sizeof(float) * TN * 3, st);
        cudaMemsetAsync(c->adam_v, 0, sizeof(float) * TN * 3, st);
    }
    cudaStreamSynchronize(st);
    mh_dev_free(dbetas);
    if (r != MH_OK) return r;
    MH_CUDA(c, cudaGetLastError());
    c->init_ready = true;
    return MH_OK;
}

extern "C" int mh_init_grads(mh_ctx* c, int32_t use_prev, int32_t use_next, void* stream) {
    API_BEGIN(c);
    if (!c->init_ready) MH_FAIL(c, MH_E_STATE, "mh_init_grads: mh_init_begin first");
    return mh_init_iter_grads(c, use_prev, use_next, (cudaStream_t)stream);
}

// ---- hot loop B ---------------------------------------------------------------------------------
__global__ void k_halo_pack(const float* __restrict__ theta, const float* __restrict__ trans, int T, int N, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * N * MH_HALO) return;
    const int side = i / (N * MH_HALO), r = i % (N * MH_HALO), n = r / MH_HALO, e = r % MH_HALO;
    const int t = side == 0 ? 0 : T - 1;
    out[i] = e < 72 ? theta[((size_t)t * N + n) * 72 + e] : trans[((size_t)t * N + n) * 3 + (e - 72)];
}

extern "C" int mh_halo_pack(mh_ctx* c, void* stream) {
    API_BEGIN(c);
    const mh_dims& d = c->d;
    k_halo_pack<<<mh_cdiv(2 * d.N * MH_HALO, 128), 128, 0, (cudaStream_t)stream>>>(c->params + c->off[MH_P_POSES_SMPL],
                                                                                  c->params + c->off[MH_P_POSES_T], d.T, d.N, c->halo_send);
    MH_LAUNCHED(c);
    return MH_OK;
}

extern "C" int mh_fit_grads(mh_ctx* c, int32_t use_prev, int32_t use_next, void* stream) {
    API_BEGIN(c);
    if (!c->model_set || !c->camera_set || !c->coefs_set || !c->ingested)
        MH_FAIL(c, MH_E_STATE, "mh_fit_grads: set model, camera, coefficients and ingest the frames first");
    cudaStream_t st = (cudaStream_t)stream;
    const mh_dims& d = c->d;
    if (d.t0 == 0) use_prev = 0;
    if (d.t0 + d.T == d.T_total) use_next = 0;
    cudaEvent_t* ev = nullptr;
    if (c->timing) {
        ev = c->events + (size_t)(c->timing_iter % MH_TIMING_RING) * MH_TIMING_EVENTS;
        c->timing_iter++;
    }
#define MH_MARK(k) do { if (ev) MH_CUDA(c, cudaEventRecord(ev[k], st)); } while (0)
    MH_MARK(0);
    MH_CUDA(c, cudaMemsetAsync(c->grads, 0, sizeof(float) * (c->n_params + MH_L_COUNT), st));
    MH_TRY(mh_loss_begin(c, st));
    MH_TRY(mh_terms_gather(c, use_prev, use_next, st));
    MhSmplArgs a = {c->params + c->off[MH_P_BETAS], d.N, 0, c->theta_all, c->trans_all, c->params + c->off[MH_P_XSCALE], c->nb, d.N,
                    c->vshaped, c->Jrest, c->A, c->pf, c->vposed, c->verts, c->j17, c->lowidx};
    MH_TRY(mh_smpl_forward_run(c, a, st));
    MH_MARK(1);
    MH_TRY(mh_terms_pre_raster(c, use_prev, use_next, st));
    MH_MARK(2);
    if (c->c.depth != 0.f || c->c.silhouette != 0.f) {
        MH_TRY(mh_render_prepass(c, st));
        MH_MARK(3);
        MH_TRY(mh_render_all(c, st));
    } else {
        MH_MARK(3);
    }
    MH_MARK(4);
    MH_TRY(mh_smpl_backward_all(c, st));
    MH_MARK(5);
    MH_TRY(mh_terms_post(c, st));
    MH_MARK(6);
#undef MH_MARK
    return MH_OK;
}

// Stage timing with CUDA events on the caller's stream (bench.py's roofline): a ring of the last MH_TIMING_RING cycles.
extern "C" int mh_set_timing(mh_ctx* c, int32_t on) {
    API_BEGIN(c);
    if (on && !c->events) {
        c->events = new cudaEvent_t[(size_t)MH_TIMING_RING * MH_TIMING_EVENTS];
        for (int i = 0; i < MH_TIMING_RING * MH_TIMING_EVENTS; ++i) MH_CUDA(c, cudaEventCreate(&c->events[i]));
    }
    c->timing = on != 0;
    c->timing_iter = 0;
    return MH_OK;
}

// out: (n_cycles, MH_TIMING_STAGES) milliseconds of the most recent cycles, oldest first; *n_cycles in/out. Blocking.
extern "C" int mh_read_timing(mh_ctx* c, float* out, int32_t* n_cycles) {
    API_BEGIN(c);
    if (!c->events || !out || !n_cycles) MH_FAIL(c, MH_E_STATE, "mh_read_timing: timing was not enabled");
    MH_CUDA(c, cudaDeviceSynchronize());
    const int64_t have = std::min<int64_t>(c->timing_iter, MH_TIMING_RING);
    const int n = (int)std::min<int64_t>(have, *n_cycles);
    for (int k = 0; k < n; ++k) {
        const int64_t it = c->timing_iter - n + k;
        cudaEvent_t* ev = c->events + (size_t)(it % MH_TIMING_RING) * MH_TIMING_EVENTS;
        for (int sgi = 0; sgi < MH_TIMING_STAGES; ++sgi) MH_CUDA(c, cudaEventElapsedTime(out + (size_t)k * MH_TIMING_STAGES + sgi, ev[sgi], ev[sgi + 1]));
    }
    *n_cycles = n;
    return MH_OK;
}

extern "C" int mh_read_losses(mh_ctx* c, float* out, void* stream) {
    API_BEGIN(c);
    cudaStream_t st = (cudaStream_t)stream;
    int flags[8];
    MH_CUDA(c, cudaMemcpyAsync(out, c->grads + c->n_params, sizeof(float) * MH_L_COUNT, cudaMemcpyDeviceToHost, st));
    MH_CUDA(c, cudaMemcpyAsync(flags, c->devflags, sizeof(flags), cudaMemcpyDeviceToHost, st));
    MH_CUDA(c, cudaStreamSynchronize(st));
    if (flags[1]) MH_FAIL(c, MH_E_CAPACITY, "render: a body exceeded the raster tile / bin capacity (%d)", flags[1]);
    return MH_OK;
}

// ---- filters ------------------------------------------------------------------------------------
extern "C" int mh_refresh_filters(mh_ctx* c, float mc1, float b1, float mc2, float b2, float frame_rate, int32_t first, void* stream) {
    API_BEGIN(c);
    MH_TRY(mh_filter_run(c, mc1, b1, mc2, b2, frame_rate, first, (cudaStream_t)stream));
    c->has_filters = true;
    return MH_OK;
}

extern "C" int mh_refresh_filters_flag(mh_ctx* c, int32_t on) {
    API_BEGIN(c);
    c->has_filters = on != 0;
    return MH_OK;
}

extern "C" int mh_clear_filters(mh_ctx* c) {
    API_BEGIN(c);
    c->has_filters = false;
    return MH_OK;
}

// ---- scene depths -------------------------------------------------------------------------------
__global__ void k_scene_depths(const float* __restrict__ depth, const float* __restrict__ zmin_lin, const float* __restrict__ zmax_lin,
                               int64_t HW, float* __restrict__ out) {
    const int t = blockIdx.y;
    // min_z = softplus(zmin_lin) ; max_z = min_z + 1 + softplus(zmax_lin)   (optimizer.py:683-688, transforms.py:296)
    const float minz = logf(1.0f + expf(zmin_lin[t]));
    const float maxz = minz + 1.0f + logf(1.0f + expf(zmax_lin[t]));
    const float ia = __fdiv_rn(1.0f, minz), b = __fdiv_rn(1.0f, maxz);
    const float a = __fsub_rn(ia, b);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)t * HW + i] = __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(depth[(int64_t)t * HW + i], a), b));     // optimizer.py:425-426
}

extern "C" int mh_scene_depths(mh_ctx* c, int32_t t0, int32_t count, float* out_host) {
    API_BEGIN(c);
    const mh_dims& d = c->d;
    if (t0 < 0 || count < 1 || t0 + count > d.T || !out_host) MH_FAIL(c, MH_E_ARG, "mh_scene_depths: bad range");
    const int64_t HW = (int64_t)d.H * d.W;
    float* tmp;
    MH_CUDA(c, mh_dev_alloc((void**)&tmp, sizeof(float) * count * HW));
    k_scene_depths<<<dim3(mh_cdiv(HW, 1024), count), 256>>>(c->depth + (int64_t)t0 * HW, c->params + c->off[MH_P_ZMIN_LIN] + t0,
                                                            c->params + c->off[MH_P_ZMAX_LIN] + t0, HW, tmp);
    c->launches++;
    cudaError_t e = cudaMemcpy(out_host, tmp, sizeof(float) * count * HW, cudaMemcpyDeviceToHost);
    mh_dev_free(tmp);
    MH_CUDA(c, e);
    return MH_OK;
}

// ---- debugging / synthesis ------------------------------------------------------------------------
extern "C" int mh_debug_render(mh_ctx* c, int32_t t, int32_t n, float* zbuf_host, float* alpha_host) {
    API_BEGIN(c);
    const mh_dims& d = c->d;
    if (t < 0 || t >= d.T || n < 0 || n >= d.N) MH_FAIL(c, MH_E_ARG, "mh_debug_render: bad index");
    const int64_t HW = (int64_t)d.H * d.W;
    float* tmp;
    MH_CUDA(c, mh_dev_alloc((void**)&tmp, sizeof(float) * 2 * HW));
    int r = mh_render_debug(c, t, n, tmp, tmp + HW, 0);
    if (r == MH_OK) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess && zbuf_host) e = cudaMemcpy(zbuf_host, tmp, sizeof(float) * HW, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && alpha_host) e = cudaMemcpy(alpha_host, tmp + HW, sizeof(float) * HW, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "mh_debug_render: %s", cudaGetErrorString(e)); r = MH_E_CUDA; }
    }
    mh_dev_free(tmp);
    return r;
}

extern "C" int mh_forward_only(mh_ctx* c, void* stream) {
    API_BEGIN(c);
    cudaStream_t st = (cudaStream_t)stream;
    const mh_dims& d = c->d;
    MH_TRY(mh_terms_gather(c, 0, 0, st));
    MhSmplArgs a = {c->params + c->off[MH_P_BETAS], d.N, 0, c->theta_all, c->trans_all, c->params + c->off[MH_P_XSCALE], c->nb, d.N,
                    c->vshaped, c->Jrest, c->A, c->pf, c->vposed, c->verts, c->j17, c->lowidx};
    return mh_smpl_forward_run(c, a, st);
}

extern "C" int mh_synth_planes(mh_ctx* c, float y_ground, float z_wall, void* stream) {
    API_BEGIN(c);
    if (!c->model_set || !c->camera_set || !c->coefs_set) MH_FAIL(c, MH_E_STATE, "mh_synth_planes: set model, camera and coefficients first");
    cudaStream_t st = (cudaStream_t)stream;
    MH_TRY(mh_forward_only(c, st));
    MH_TRY(mh_render_synth(c, y_ground, z_wall, st));
    return MH_OK;
}

extern "C" int mh_read_planes(mh_ctx* c, int32_t t0, int32_t count, float* depths_host, float* seg_host) {
    API_BEGIN(c);
    const mh_dims& d = c->d;
    if (t0 < 0 || count < 1 || t0 + count > d.T) MH_FAIL(c, MH_E_ARG, "mh_read_planes: bad range");
    const int64_t HW = (int64_t)d.H * d.W;
    MH_CUDA(c, cudaDeviceSynchronize());
    if (depths_host) MH_CUDA(c, cudaMemcpy(depths_host, c->depth + (int64_t)t0 * HW, sizeof(float) * count * HW, cudaMemcpyDeviceToHost));
    if (seg_host) {
        float* tmp;
        MH_CUDA(c, mh_dev_alloc((void**)&tmp, sizeof(float) * d.N * HW));
        int r = MH_OK;
        for (int t = t0; t < t0 + count && r == MH_OK; ++t) {
            r = mh_expand_planes(c, t, tmp, 0);
            if (r == MH_OK && cudaMemcpy(seg_host + (int64_t)(t - t0) * d.N * HW, tmp, sizeof(float) * d.N * HW, cudaMemcpyDeviceToHost) != cudaSuccess) {
                snprintf(c->err, sizeof(c->err), "mh_read_planes: copy failed");
                r = MH_E_CUDA;
            }
        }
        mh_dev_free(tmp);
        return r;
    }
    return MH_OK;
}
