// Internal context of libmhopt.so (not part of the C ABI; see include/mhopt.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mhopt.h"
#include "mh_math.cuh"

#define MH_LD3V 20672          // padded row stride (floats) of every (*, 3V) matrix: 3*6890 = 20670 -> multiple of 32
#define MH_KPF 192             // padded pose-feature length (189 live rows)
#define MH_NEXT 208            // rows of the extended basis: 0..191 posedirs (189 live), 192..201 shapedirs, pad
#define MH_KSPLIT 19            // split-K factor of the backward contraction (20672 = 19 * 1088)
#define MH_MAXN 32             // persons per frame (bit masks are 32-bit)
#define MH_NUM_SMS_FALLBACK 148

struct MhRenderScratch;        // mh_render.cu

struct mh_ctx {
    mh_dims d;
    int Ts;                    // frame slots = T + 2 (slot 0: halo of the previous rank, slot T+1: halo of the next)
    int nb;                    // bodies = Ts * N (slot-major)
    int num_sms;
    char err[512];
    int64_t launches;
    bool model_set, camera_set, ingested, optim_scale;
    // ---- model ----
    float* pext;               // (MH_NEXT, MH_LD3V) extended basis
    float* vtemplate;          // (MH_LD3V)
    float* Jt;                 // (24,3)   J_regressor . v_template
    float* Js;                 // (24,3,10) J_regressor . shapedirs
    int KW;                    // skinning weights kept per vertex
    uint8_t* wj;               // (V, KW)
    float* ww;                 // (V, KW)
    int* jptr; int* jvert; float* jw; int jnnz;      // CSC of the skinning weights (joint -> vertices)
    int* rptr; int* rvert; float* rw; int rnnz;      // CSR of the 17-joint regressor
    int32_t* faces;            // (F,3)
    // ---- camera / coefficients ----
    float K[9], Kndc[16], Kd[5];
    bool has_kd;
    mh_coefs c;
    // ---- frame data (reference layout: one f32 plane per person + one f32 disparity plane per frame) ----
    float* depth;              // (T, H*W)
    float* seg;                // (T, N, H*W)
    uint32_t* ebits;           // (T, H*W)   bit n = erode5x5(seg[t,n])   (optimizer.py:306-309)
    uint8_t* rankplane;        // (T, H*W)   first depth-order position whose mask covers the pixel (255 none)
    float* pose2d;             // (T, N, 17, 3)
    float* theta_ref;          // (T, N, 72)
    float* valid;              // (T, N)
    float* maskarea;           // (T, N)
    uint8_t* pose2d_valid;     // (T, N)  >= 2 joints over the confidence threshold (optimizer.py:404-405)
    uint8_t* mask_valid;       // (T, N)  mask area >= 0.5 % of the image (optimizer.py:407-409)
    // ---- parameters: one flat buffer [poses_T | poses_smpl | zmin | zmax | betas | xscale] ----
    int64_t off[MH_P_COUNT];   // float offsets of the six leaves (+ betas_ref kept outside)
    int64_t n_params;
    float* params; float* grads; float* sqavg; float* mom;   // grads has 16 extra floats (loss block) at the end
    float* betas_ref;          // (N,10)
    float* halo_send; float* halo_recv;                     // (2, N, 75)
    // ---- init stage (hot loop A) ----
    float* init_j17;           // (T, N, 17, 3) local regressed joints for the per-frame ROMP betas
    float* init_vis;           // (T, N, 17)
    float* adam_m; float* adam_v;
    // ---- per-iteration scratch ----
    float* theta_all;          // (nb, 72)
    float* trans_all;          // (nb, 3)
    float* vshaped;            // (max(N, init bodies), MH_LD3V)
    int64_t vshaped_rows;
    float* Jrest;              // (rows, 72)
    float* A;                  // (nb, 24, 12)
    float* pf;                 // (nb, MH_KPF)
    float* vposed;             // (nb, MH_LD3V)
    float* verts;              // (nb, MH_LD3V)   absolute vertices  scale * v + T   (optimizer.py:702)
    float* dverts;             // (nb, MH_LD3V)   dL/dverts, then reused for dL/dv_posed
    float* filtered;           // (nb, MH_LD3V)   One-Euro filtered vertices (optimizer.py:390-392)
    bool has_filters;
    float* j17;                // (nb, 17, 3)
    int* lowidx;               // (nb)  argmax_v y   (optimizer.py:487)
    float* dA;                 // (nb, 24, 12)
    float* gT;                 // (nb, 4): sum dV (3), sum <dV, v_local>
    float* dpf_part;           // (MH_KSPLIT, nb, MH_NEXT)
    // order / silhouette prepass
    int* order;                // (T, N) persons sorted near -> far
    float* sumM;               // (T, N)  sum over the image of (1 - acc) per order position
    float* silbase;            // (T, N)  sum over the image of ((1 - acc) * seg)^2 for alpha = 0
    // per person-frame raster outputs
    float* pfout;              // (T*N, 8): depth loss, silhouette loss, G_{1/zmin}, G_{1/zmax}, S, ...
    // contact
    float* scene; int64_t M;   // (M,3)
    float* contact;            // (T*N, 4): cdv, in-contact flag, ...
    float* footacc;            // (num local batches, 2): numerator, denominator
    int n_batches;
    // one-euro carry
    float* carry_in; float* carry_out; int64_t carry_floats;
    float* transfilt;          // (Ts*N*3) filtered translations (only a not-None flag upstream)
    MhRenderScratch* rs;
    int64_t adam_n;
};

#define MH_FAIL(ctx, code, ...) do { snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); return (code); } while (0)
#define MH_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    return MH_E_CUDA; } } while (0)
#define MH_LAUNCHED(ctx) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    return MH_E_CUDA; } } while (0)

static inline int mh_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// stage entry points implemented in the .cu files (all enqueue on `st`)
int mh_smpl_forward_all(mh_ctx* c, int first_body, int n_bodies, bool per_body_shape, const float* betas_dev,
                        const float* theta_dev, const float* trans_dev /*or null*/, const float* xscale_dev /*or null*/,
                        float* verts_out, float* j17_out, int* lowidx_out, cudaStream_t st);
int mh_smpl_backward_all(mh_ctx* c, cudaStream_t st);
int mh_terms_forward(mh_ctx* c, int use_prev, int use_next, cudaStream_t st);
int mh_terms_finalize(mh_ctx* c, cudaStream_t st);
int mh_render_alloc(mh_ctx* c);
void mh_render_free(mh_ctx* c);
int mh_render_prepass(mh_ctx* c, cudaStream_t st);
int mh_render_all(mh_ctx* c, cudaStream_t st);
int mh_render_debug(mh_ctx* c, int t, int n, float* zbuf_dev, float* alpha_dev, float blur_d, float blur_s, cudaStream_t st);
int mh_ingest_derive(mh_ctx* c, cudaStream_t st);
