// Internal context of libmhopt.so (not part of the C ABI; see include/mhopt.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/mhopt.h"
#include "mh_math.cuh"

#define MH_V 6890
#define MH_F 13776
#define MH_LD3V 20672          // padded row stride (floats) of every (*, 3V) matrix: 3*6890 = 20670 -> multiple of 32
#define MH_KPF 192             // padded pose-feature length (189 live rows)
#define MH_NEXT 208            // rows of the extended basis: 0..191 posedirs (189 live), 192..201 shapedirs, pad
#define MH_KSPLIT 19           // split-K factor of the backward contraction (20672 = 19 * 1088)
#define MH_MAXN 32             // persons per frame (bit masks are 32-bit)
#define MH_KNN 32              // scene points averaged by the contact term (optimizer.py:494)
#define MH_TIMING_RING 64
#define MH_TIMING_STAGES 6      // smpl forward | pre-raster terms | order prepass | render | smpl backward | post terms
#define MH_TIMING_EVENTS 7
#define MH_HALO 75             // floats per person in a halo frame: theta (72) + T (3)

// per person-frame outputs of the render stage (floats)
enum { PF_S = 0, PF_A, PF_C, PF_GIZMIN, PF_GIZMAX, PF_SIL, PF_CNTIN, PF_DEPTHLOSS, PF_COUNT = 8 };

struct MhRenderScratch;        // mh_render.cu

struct mh_ctx {
    mh_dims d;
    int Ts;                    // frame slots = T + 2 (slot 0: halo of the previous rank, slot T+1: halo of the next)
    int nb;                    // bodies = Ts * N (slot-major)
    int num_sms;
    char err[512];
    int64_t launches;
    bool model_set, camera_set, coefs_set, ingested, optim_scale, has_filters, init_ready;
    std::vector<void*> allocs;
    // ---- model ----
    float* pext;               // (MH_NEXT, MH_LD3V) extended basis
    float* pextF;              // the basis pre-split (hi | lo TF32) in the UMMA tile layout of the forward contraction (mh_gemm_tc.cu)
    float* pextB;              // ... of the backward contraction
    float* vtemplate;          // (MH_LD3V)
    float* Jt;                 // (24,3)   J_regressor . v_template
    float* Js;                 // (24,3,10) J_regressor . shapedirs
    int KW;                    // skinning weights kept per vertex
    uint8_t* wj;               // (V, KW)
    float* ww;                 // (V, KW)
    int* jptr; int* jvert; float* jw; int jnnz;      // CSC of the skinning weights (joint -> vertices)
    int* rptr; int* rvert; float* rw; int rnnz;      // CSR of the 17-joint regressor (joint -> vertices)
    int* cptr; int* cjoint; float* cw;               // CSC of the 17-joint regressor (vertex -> joints)
    int32_t* faces;            // (F,3)
    // ---- camera / coefficients ----
    float K[9], Kndc[16], Kd[5];
    bool has_kd;
    float* pix_x;              // (W) NDC x of every pixel column centre (PyTorch3D convention)
    float* pix_y;              // (H)
    mh_coefs c;
    float w17[MH_NJR];         // pose17j_weights after normalisation (optimizer.py:127-130)
    float r17_slack[MH_NJR];   // 1 - sum of the regressor row: J17 = R17 . V + T (1 - rowsum) (0 for the shipped regressors)
    // ---- frame data: f32 disparity plane per frame + two 32-bit person bit planes per frame ----
    float* depth;              // (T, H*W)
    uint32_t* cbits;           // (T, H*W)   bit n = seg_mask[t,n] > 0                (optimizer.py:407, 475)
    uint32_t* ebits;           // (T, H*W)   bit n = erode(erode(seg_mask[t,n]))       (optimizer.py:306-309, 434)
    float* stage;              // staging for one ingest call (count, N, H*W) f32
    int64_t stage_floats;
    // host-side compaction of float32 masks (ingest): two pinned buffers of packed bit planes, each with the event of its last copy
    uint32_t* pack_buf[2]; cudaEvent_t pack_ev[2]; int64_t pack_words; int pack_idx; bool host_nonbinary;
    float* pose2d;             // (T, N, 17, 3)
    float* theta_ref;          // (T, N, 72)
    float* valid;              // (T, N)
    int* maskarea;             // (T, N)
    uint8_t* pose2d_valid;     // (T, N)  >= 2 joints over the confidence threshold (optimizer.py:404-405)
    uint8_t* mask_valid;       // (T, N)  mask area >= 0.5 % of the image (optimizer.py:407-409)
    int* devflags;             // [0] non-binary mask seen at ingest, [1] raster capacity exceeded
    // ---- parameters: one flat buffer [poses_T | poses_smpl | zmin | zmax | betas | xscale] ----
    int64_t off[MH_P_COUNT];
    int64_t cnt[MH_P_COUNT];
    int64_t n_params;          // floats in the six leaves
    float* params; float* grads; float* sqavg; float* mom;   // grads has MH_L_COUNT extra floats (loss block) at the end
    float* betas_ref;          // (N,10)
    float* halo_send; float* halo_recv;                     // (2, N, 75)
    // ---- init stage (hot loop A) ----
    float* init_j17;           // (T, N, 17, 3) local regressed joints for the per-frame ROMP betas
    float* init_vis;           // (T, N, 17)
    float* adam_m; float* adam_v;   // (T, N, 3)
    // ---- per-iteration scratch ----
    float* theta_all;          // (nb, 72)
    float* trans_all;          // (nb, 3)
    float* vshaped;            // (N, MH_LD3V)
    float* Jrest;              // (nb, 72) (N rows used by fit; per-body rows by the init / utility paths)
    float* A;                  // (nb, 24, 12)
    float* pf;                 // (nb, MH_KPF)
    float* vposed;             // (nb, MH_LD3V)
    float* verts;              // (nb, MH_LD3V)   absolute vertices  scale * v + T   (optimizer.py:702)
    float* dverts;             // (nb, MH_LD3V)   dL/dverts, then reused for dL/dv_posed
    float* filtered;           // (nb, MH_LD3V)   One-Euro filtered vertices (optimizer.py:390-392)
    float* j17;                // (nb, 17, 3)
    float* gj17;               // (nb, 17, 3)
    int* lowidx;               // (nb)  argmax_v y   (optimizer.py:487)
    float* dA;                 // (nb, 24, 12)
    float* gT;                 // (nb, 4): sum dV (3), sum <dV, v_local>
    float* lpart; int LP;      // (MH_L_COUNT, LP) loss partials of one cycle, one slot per contributor (LP = 8 T N), summed in a fixed order
    float* shared_part;        // (T*N, 12): per-body contributions to the shared-leaf gradients (10 betas, xscale), reduced in a fixed order
    float* dpf_part;           // (MH_KSPLIT + 1, nb, MH_NEXT)
    // depth order / silhouette prepass
    int* order;                // (T, N) persons sorted near -> far (optimizer.py:450)
    uint32_t* premask;         // (T, N) union of the person bits in front of position q
    int* dirty;                // (T) order changed since the last prepass
    int* rankcnt;              // (T, N+1) pixels whose first covering order position is q (N = none)
    // per person-frame raster outputs
    float* pfout;              // (T*N, PF_COUNT)
    // contact
    float* scene; int64_t M;   // (M,3)
    float* contact;            // (T*N, 4): cdv, in-contact flag
    // one-euro carry
    float* carry_in; float* carry_out; int64_t carry_floats;
    float* transfilt;          // (T*N*3) filtered translations (only a not-None flag upstream)
    MhRenderScratch* rs;
    void* scene_state;         // mh_scene.cu
    void* scene_post;          // mh_scenepost.cu
    void* knn;                 // mh_terms.cu: uniform grid over the scene cloud for the contact term
    int* scene_counts;         // mh_filter.cu: per-block point counts of the scene-cloud compaction
    void* comm;                // mh_comm.cu: the NCCL communicator this context owns (mh_set_comm), or null
    // stage timing (bench)
    cudaEvent_t* events; bool timing; int64_t timing_iter;
};

#define MH_FAIL(ctx, code, ...) do { snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); return (code); } while (0)
#define MH_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    return MH_E_CUDA; } } while (0)
#define MH_LAUNCHED(ctx) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    return MH_E_CUDA; } } while (0)
#define MH_TRY(call) do { int r_ = (call); if (r_ != MH_OK) return r_; } while (0)

static inline int mh_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// SMPL forward over `nbodies` consecutive bodies; all arrays are indexed from body 0 of the call.
struct MhSmplArgs {
    const float* betas;        // (shape_rows, 10)
    int shape_rows;            // rows of vshaped / Jrest to prepare (N, or nbodies when per_body_shape)
    int per_body_shape;        // 0: body b uses shape row b % N ; 1: body b uses row b
    const float* theta;        // (nbodies, 72)
    const float* trans;        // (nbodies, 3) or null
    const float* xscale;       // (N) or null  -> scale 1.1^x (optimizer.py:681)
    int nbodies, N;
    float* vshaped; float* Jrest; float* A; float* pf; float* vposed;
    float* verts;              // (nbodies, MH_LD3V)
    float* j17;                // (nbodies, 17, 3) or null
    int* lowidx;               // (nbodies) or null
};
int mh_smpl_forward_run(mh_ctx* c, const MhSmplArgs& a, cudaStream_t st);
int mh_gemm_fwd_tc(mh_ctx* c, const float* pf, const float* vshaped, float* vposed, int nbodies, int Npers, int per_body_shape,
                   cudaStream_t st);          // mh_gemm_tc.cu: tcgen05 / TMEM, 3 x TF32
int mh_gemm_bwd_tc(mh_ctx* c, const float* E, float* dpf_part, int M, int first_body, int nb_total, cudaStream_t st);
int mh_gemm_tc_prepare(mh_ctx* c);                                          // builds pextF / pextB from c->pext at mh_set_model
int mh_upload_floats(mh_ctx* c, float** p, const std::vector<float>& h);    // allocation owned by the context + H2D copy
int mh_alloc_floats(mh_ctx* c, float** p, int64_t n);                       // zeroed allocation owned by the context
int mh_alloc_ints(mh_ctx* c, int** p, int64_t n);
int mh_gemm_bwd_simt(mh_ctx* c, const float* E, float* dpf_part, int M, int first_body, int nb_total, cudaStream_t st);
int mh_gemm_fwd_simt(mh_ctx* c, const float* pf, const float* vshaped, float* vposed, int nbodies, int Npers, int per_body_shape, cudaStream_t st);
int mh_smpl_backward_all(mh_ctx* c, cudaStream_t st);

// mh_terms.cu
int mh_terms_gather(mh_ctx* c, int use_prev, int use_next, cudaStream_t st);
int mh_terms_pre_raster(mh_ctx* c, int use_prev, int use_next, cudaStream_t st);
int mh_terms_post(mh_ctx* c, cudaStream_t st);
// pooled device memory (mh_pool.cu): every device buffer of the library comes from / returns to these
cudaError_t mh_dev_alloc(void** p, size_t bytes);
cudaError_t mh_dev_free(void* p);
template <typename T> static inline cudaError_t mh_dev_alloc(T** p, size_t bytes) { return mh_dev_alloc((void**)p, bytes); }
int mh_knn_build(mh_ctx* c, cudaStream_t st);       // (re)build the contact-term grid for the current scene cloud
void mh_knn_free(mh_ctx* c);
int mh_loss_begin(mh_ctx* c, cudaStream_t st);      // zero the loss partials of the cycle
int mh_loss_reduce(mh_ctx* c, cudaStream_t st);     // losses[slot] += fixed-order sum of the slot's partials
int mh_init_iter_grads(mh_ctx* c, int use_prev, int use_next, cudaStream_t st);
// mh_render.cu
int mh_render_alloc(mh_ctx* c);
void mh_render_free(mh_ctx* c);
int mh_render_prepass(mh_ctx* c, cudaStream_t st);
int mh_render_all(mh_ctx* c, cudaStream_t st);
int mh_render_debug(mh_ctx* c, int t, int n, float* zbuf_dev, float* alpha_dev, cudaStream_t st);
int mh_render_planes(mh_ctx* c, int t, int n, float* zbuf_dev, float* alpha_dev, float blur_d, float blur_s, cudaStream_t st);
int mh_render_synth(mh_ctx* c, float y_ground, float z_wall, cudaStream_t st);
// mh_filter.cu
int mh_filter_run(mh_ctx* c, float mc1, float b1, float mc2, float b2, float frame_rate, int first, cudaStream_t st);
int mh_scene_from_depth(mh_ctx* c, const float* depth_dev, const uint8_t* mask_dev, cudaStream_t st);
int mh_ingest_compact(mh_ctx* c, int t0, int count, int seg_is_u8, cudaStream_t st);
int mh_ingest_derive(mh_ctx* c, cudaStream_t st);
int mh_expand_planes(mh_ctx* c, int t, float* seg_dev, cudaStream_t st);
// mh_comm.cu
void mh_comm_free(mh_ctx* c);
// mh_scenepost.cu
int mh_scene_postprocess_dev(mh_ctx* c, const float* depth_dev, const uint8_t* mask_dev, int use_bilateral, int fillin_ksize, cudaStream_t st);
const float* mh_scene_post_result(mh_ctx* c);
void mh_scenepost_free(mh_ctx* c);
// mh_scene.cu
void mh_scene_free(mh_ctx* c);
int mh_scene_views(mh_ctx* c, int which, void** ptr, int64_t* n);
