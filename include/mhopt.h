/* libmhopt.so -- C ABI of the B200-native scene-aware multi-human SMPL optimiser.
 *
 * The reference (dluvizon/scene-aware-3d-multi-human) has no FFI layer: the
 * drop-in boundary is the Python class `SMPLDepthSequenceOptimizer`
 * (mhmocap/optimizer.py:146-770) that `Predictor` drives (mhmocap/predict.py:290-306,
 * 332-344).  The Python clone in `scene-aware-3d-multi-human_b200/optimizer.py` keeps
 * that class's signatures and binds the entry points below with ctypes
 * (INTEGRATION.md shows the binding).  Each entry point cites the reference
 * code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MH_E_* code; nothing
 *     throws or aborts across the boundary; `mh_last_error` gives the text;
 *   - plain pointers and sizes only: no torch / pybind types;
 *   - pointers named `*_host` are HOST pointers (pinned or pageable) that the
 *     library copies from/to with cudaMemcpyAsync on `stream`; pointers named
 *     `*_dev` are device pointers owned by the library (zero-copy views for the
 *     caller's NCCL plumbing) that stay valid until `mh_destroy`;
 *   - `stream` is a `cudaStream_t` passed as void* (NULL = legacy default
 *     stream); work is enqueued on it, the library never synchronises except
 *     in the explicitly blocking getters (`mh_get_*`, `mh_read_losses`);
 *   - one context per GPU / rank; NOT thread-safe, the caller serialises;
 *   - all floating point data is float32, indices int32, row-major.
 */
#ifndef MHOPT_H_
#define MHOPT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MH_OK 0
#define MH_E_ARG (-1)      /* bad argument / shape */
#define MH_E_CUDA (-2)     /* CUDA runtime error */
#define MH_E_STATE (-3)    /* call made in the wrong order (e.g. fit before ingest) */
#define MH_E_CAPACITY (-4) /* an internal capacity (raster bins, scene cloud) was exceeded */

typedef struct mh_ctx mh_ctx;

typedef struct mh_dims {
    int32_t T;        /* frames owned by this rank (contiguous range) */
    int32_t N;        /* persons (<= 32) */
    int32_t H, W;     /* image rows, columns */
    int32_t V, F;     /* SMPL vertices (6890), faces (13776) */
    int32_t B;        /* frames per batch segment: SEMANTIC (optimizer.py:512-518, 526, 531-542) */
    int32_t device;   /* CUDA ordinal */
    int32_t rank, world;
    int32_t t0;       /* global index of this rank's first frame */
    int32_t T_total;  /* frames of the whole sequence */
    int64_t M_max;    /* capacity of the scene point cloud */
} mh_dims;

/* SMPL model buffers as `SMPL.__init__` registers them (mhmocap/smpl.py:201-275). HOST pointers. */
typedef struct mh_model {
    const float* v_template;    /* (V,3) */
    const float* shapedirs;     /* (V,3,10) */
    const float* posedirs;      /* (207, 3V) */
    const float* J_regressor;   /* (24,V) */
    const float* lbs_weights;   /* (V,24) */
    const int32_t* parents;     /* (24) parents[0] = -1 */
    const int32_t* faces;       /* (F,3) */
    const float* reg17;         /* (17,V) J_regressor_alphapose (smpl.py:249-252) */
} mh_model;

/* Loss coefficients: SMPLDepthSequenceOptimizer.__init__ (optimizer.py:159-169). */
typedef struct mh_coefs {
    float proj2d, depth, silhouette, reg_velocity, reg_verts_filter, reg_poses, reg_scales, reg_contact,
          reg_foot_sliding;
    float joint_confidence_thr;   /* 0.5 */
    float eps;                    /* 1e-3 */
} mh_coefs;

/* Optimised leaves (optimizer.py:343-351) and read-only per-frame data, for mh_set_param / mh_get_param. */
enum {
    MH_P_POSES_T = 0,    /* (T,N,3)   optimizer.py:291 */
    MH_P_POSES_SMPL = 1, /* (T,N,72)  optimizer.py:295 */
    MH_P_BETAS = 2,      /* (N,10)    optimizer.py:297 */
    MH_P_ZMIN_LIN = 3,   /* (T)       optimizer.py:302 */
    MH_P_ZMAX_LIN = 4,   /* (T)       optimizer.py:303 */
    MH_P_XSCALE = 5,     /* (N)       optimizer.py:284 */
    MH_P_BETAS_REF = 6,  /* (N,10)    optimizer.py:298 (frozen copy used by the shape prior) */
    MH_P_COUNT = 7
};

/* Entries of the 16-float loss block written by mh_fit_grads (names of optim_log, optimizer.py:546-554, 592-593). */
enum {
    MH_L_POSE2D = 0, MH_L_DEPTH = 1, MH_L_SILHOUETTE = 2, MH_L_REF_POSES = 3, MH_L_SCALE = 4, MH_L_CONTACT = 5,
    MH_L_FOOT = 6, MH_L_VEL = 7, MH_L_FILTER_VERTS = 8, MH_L_INIT_2D = 9, MH_L_COUNT = 16
};

/* Zero-copy device views for the caller's collective plumbing (torch.distributed / NCCL). */
enum {
    MH_BUF_SHARED = 0,     /* [g_betas (N*10) | g_xscale (N) | losses (16)]: all-reduce(sum) across ranks */
    MH_BUF_HALO_SEND = 1,  /* [first frame theta,T (N*75) | last frame theta,T (N*75)] */
    MH_BUF_HALO_RECV = 2,  /* [prev rank's last frame (N*75) | next rank's first frame (N*75)] */
    MH_BUF_CARRY_OUT = 3,  /* One-Euro state after this rank's last frame: x_prev, dx_prev for T and verts */
    MH_BUF_CARRY_IN = 4,   /* same layout, state handed over by the previous rank */
    MH_BUF_GRADS = 5,      /* the whole flat gradient buffer (tests) */
    MH_BUF_VERTS = 6,      /* (T+2, N, 20672) posed vertices of the last forward incl. halo slots; rows padded to 20672 floats */
    MH_BUF_FILTERED = 7,   /* (T+2, N, 20672) filtered vertices (optimizer.py:390-392) incl. halo slots; rows padded to 20672 floats */
    MH_BUF_PARAMS = 8,     /* flat parameter buffer [poses_T | poses_smpl | zmin_lin | zmax_lin | betas | xscale] */
    MH_BUF_MEDIAN_HIST = 9, /* (48, H*W) per-pass digit histograms of the scene median: all-reduce(SUM) between passes */
    MH_BUF_MEDIAN_AUX = 10  /* (6, H*W) last pass: planes 0-2 all-reduce(SUM), planes 3-5 all-reduce(MIN) */
};

/* ---- lifetime -------------------------------------------------------------------------------- */
int mh_create(mh_ctx** out, const mh_dims* dims);
void mh_destroy(mh_ctx* ctx);
const char* mh_last_error(const mh_ctx* ctx);
const char* mh_version(void);

/* frames per dataloader batch: SEMANTIC (optimizer.py:512-518, 526, 531-542); may be changed before fit() */
int mh_set_batch(mh_ctx* ctx, int32_t B);

/* ---- constant inputs ------------------------------------------------------------------------- */
/* SMPLOptimizerBase.__init__ (optimizer.py:64-75): uploads the model and builds the sparse forms. */
int mh_set_model(mh_ctx* ctx, const mh_model* model_host);
/* cam_K (optimizer.py:193-200), the PyTorch3D NDC matrix of transforms.py:222-255 (optimizer.py:206), Kd (:200) */
int mh_set_camera(mh_ctx* ctx, const float K3x3_host[9], const float Kndc4x4_host[16], const float* Kd5_host_or_null);
int mh_set_coefs(mh_ctx* ctx, const mh_coefs* coefs);
/* pose17j_weights AFTER the normalisation of optimizer.py:127-130 (17 floats; default all 1) */
int mh_set_joint_weights(mh_ctx* ctx, const float* w17_host);
/* 0: xscale_factor is fixed (scale_factor given to init_optimized_variables, optimizer.py:279-282, 350-353) */
int mh_set_optimize_scale(mh_ctx* ctx, int32_t on);

/* One dataloader batch (optimizer.py:394-400; keys of datautils.py:630-641), frames
 * [t_local0, t_local0 + count): depths (count,H,W), seg_mask (count,N,H,W), pose2d (count,N,17,3),
 * poses_smpl reference (count,N,72), valid_smpl (count,N) already thresholded > 0.7 (optimizer.py:299).
 * HOST pointers; the copy is asynchronous on `stream` (keep the buffers alive until it completes).
 * depths_host / seg_host may be NULL (planes already on the device, e.g. after mh_synth_planes). */
int mh_ingest_frames(mh_ctx* ctx, int32_t t_local0, int32_t count, const float* depths_host,
                     const float* seg_host, const float* pose2d_host, const float* theta_ref_host,
                     const float* valid_host, void* stream);
/* The same with the instance masks as uint8 / bool {0,1} (count,N,H,W): a quarter of the host->device bytes of the reference
 * dataset's float32 masks (utils.py:329-331 delivers {0.,1.}); extension for callers that can hand the masks over compactly. */
int mh_ingest_frames_u8(mh_ctx* ctx, int32_t t_local0, int32_t count, const float* depths_host,
                        const uint8_t* seg_host, const float* pose2d_host, const float* theta_ref_host,
                        const float* valid_host, void* stream);
/* Derived constants after the last ingest: eroded masks (optimizer.py:306-309, 434-435), mask areas and
 * validity flags (optimizer.py:404-409). */
int mh_finalize_ingest(mh_ctx* ctx, void* stream);

/* update_scene_pointcloud (optimizer.py:605-616): replicated read-only scene cloud (M,3); M = 0 disables
 * the contact / foot-sliding terms (optimizer.py:485). */
int mh_set_scene(mh_ctx* ctx, const float* pcd_xyz_host, int64_t M, void* stream);
/* same from a device-resident scene depth map + mask: inverse projection of the pixel centres (optimizer.py:609-613) */
int mh_set_scene_from_depth(mh_ctx* ctx, const float* depth_host, const uint8_t* mask_host, void* stream);

/* ---- parameters ------------------------------------------------------------------------------ */
int mh_set_param(mh_ctx* ctx, int which, const float* src_host, int64_t count, void* stream);
int mh_get_param(mh_ctx* ctx, int which, float* dst_host, int64_t count);          /* blocking */
int mh_get_grad(mh_ctx* ctx, int which, float* dst_host, int64_t count);           /* blocking; MH_P_* leaves */
int mh_device_view(mh_ctx* ctx, int which_buf, void** dev_ptr, int64_t* n_floats);
/* resets the RMSprop state and the step counter (start of fit(), optimizer.py:355-356) */
int mh_reset_optimizer(mh_ctx* ctx, void* stream);

/* ---- SMPL forward utility: SMPL.forward (smpl.py:297-399) ------------------------------------- */
/* betas (nb,10), theta (nb,72) -> verts (nb,V,3) and joints_alphapose (nb,17,3); HOST pointers, blocking;
 * either output may be NULL. */
int mh_smpl_forward(mh_ctx* ctx, const float* betas_host, const float* theta_host, int64_t nb,
                    float* verts_host, float* joints17_host);
/* evaluation (mhmocap/evaluate.py:222-229, smpl.py:376-389): SMPL forward of nbodies (betas (nb,10), theta (nb,72)), then
 * joints (nb,J,3) = regressor (J,6890, dense, HOST) . local vertices (no scale, no translation).  HOST pointers, blocking. */
int mh_smpl_regress(mh_ctx* ctx, const float* betas, const float* theta, int64_t nbodies, const float* regressor, int32_t J, float* joints);

/* ---- hot loop A: __init_global_poses (optimizer.py:710-770) ------------------------------------ */
/* pose2d (T,N,17,3), per-frame ROMP theta (T,N,72) / betas (T,N,10): evaluates the (constant) regressed
 * joints once, sets poses_T = (0,0,1) and resets the Adam state. */
int mh_init_begin(mh_ctx* ctx, const float* pose2d_host, const float* theta_host, const float* betas_host,
                  float joints_thr, void* stream);
/* one Adam iteration on poses_T (lr, betas (0.5,0.5), eps 1e-6; bias-corrected); writes loss_2d into
 * losses[MH_L_INIT_2D] (sum of squares over local frames; the caller divides by T_total*N*17*2). */
int mh_init_grads(mh_ctx* ctx, int32_t use_halo_prev, int32_t use_halo_next, void* stream);
int mh_init_update(mh_ctx* ctx, float lr, int32_t step_1based, void* stream);

/* ---- hot loop B: one cycle of fit() (optimizer.py:375-593) ------------------------------------- */
/* packs this rank's boundary frames into MH_BUF_HALO_SEND (temporal terms, optimizer.py:560-575) */
int mh_halo_pack(mh_ctx* ctx, void* stream);
/* forward + analytic backward of every term; gradients of the per-frame leaves are final, the shared block
 * (MH_BUF_SHARED) holds this rank's partial sums.  `use_halo_prev/next`: MH_BUF_HALO_RECV holds valid data. */
int mh_fit_grads(mh_ctx* ctx, int32_t use_halo_prev, int32_t use_halo_next, void* stream);
/* torch.optim.RMSprop(lr, alpha .5, eps 1e-8, momentum .9) step on all leaves (optimizer.py:355, 586) */
int mh_fit_update(mh_ctx* ctx, float lr, void* stream);
/* blocking read of the 16-float loss block (after the caller's all-reduce) */
int mh_read_losses(mh_ctx* ctx, float* out16_host, void* stream);

/* ---- One-Euro refresh (optimizer.py:383-392, 664-675) ------------------------------------------ */
/* scans this rank's frames on the device; `first` != 0 starts a new filter at local frame 0, otherwise the
 * state in MH_BUF_CARRY_IN is continued; leaves the end state in MH_BUF_CARRY_OUT. */
int mh_refresh_filters(mh_ctx* ctx, float min_cutoff1, float beta1, float min_cutoff2, float beta2,
                       float frame_rate, int32_t first, void* stream);
int mh_clear_filters(mh_ctx* ctx);
/* marks MH_BUF_FILTERED as valid / invalid when the caller filled it through the device view (teacher-forced tests) */
int mh_refresh_filters_flag(mh_ctx* ctx, int32_t on);
/* SMPLDepthSequenceOptimizer.one_euro_filter (optimizer.py:664-675) of a HOST array (T, row_elems); blocking */
int mh_one_euro_filter(mh_ctx* ctx, const float* x_host, float* y_host, int32_t T, int64_t row_elems, float min_cutoff,
                       float beta, float frame_rate);

/* ---- scene-geometry inputs (optimizer.py:425-426, 578-584) ------------------------------------- */
/* per-frame scene depth 1 / target_disp for the current parameters, into a HOST buffer (count,H,W); blocking */
int mh_scene_depths(mh_ctx* ctx, int32_t t_local0, int32_t count, float* out_host);

/* SMPL forward of every local person-frame with the current parameters into MH_BUF_VERTS (no losses) */
int mh_forward_only(mh_ctx* ctx, void* stream);

/* ---- scene geometry: masked temporal median on the device (fhsog.py:180-202, optimizer.py:578-582) -------------------- */
/* background masks (count,H,W) u8 and, optionally, RGB frames (count,H,W,3) u8 of local frames [t_local0, t_local0+count) */
int mh_scene_set_back(mh_ctx* ctx, int32_t t_local0, int32_t count, const uint8_t* backmask_host, const uint8_t* images_host_or_null,
                      void* stream);
/* one pass of the exact radix selection; which: 0 = depth 1/target_disp of the CURRENT parameters (10 passes), 1 = image
 * (4 passes).  Between passes the caller all-reduces MH_BUF_MEDIAN_HIST (SUM) when the frames are sharded; after the last pass
 * MH_BUF_MEDIAN_AUX (planes 0-2 SUM, planes 3-5 MIN). */
int mh_scene_median_pass(mh_ctx* ctx, int32_t which, int32_t pass, void* stream);
/* results to HOST: depth (H,W) f32 + mask (H,W) u8 (which = 0) or image (H,W,3) u8 (which = 1); blocking */
/* ---- library-owned communicator and fused cycles (mh_set_comm; 8e) -------------------------------------------------
 * The reference is single-process (predict.py:267-271); these serve this build's frame sharding.  NCCL is bound at run time
 * (dlopen): without it mh_comm_unique_id / mh_set_comm return MH_E_STATE and the caller exchanges the buffers of mh_device_view
 * itself (optimizer.py does so through torch.distributed). */
/* 128-byte NCCL unique id, produced on ONE rank and handed to every rank's mh_set_comm by the caller */
int mh_comm_unique_id(mh_ctx* ctx, uint8_t* out128);
/* collective over the mh_dims.world ranks of `ctx` (its rank / world / device are used): creates an ncclComm_t that outlives the
 * context -- ncclCommInitRank takes seconds on an 8-GPU node, a process creates it once for all the sequences it fits */
int mh_comm_create(mh_ctx* ctx, const uint8_t* unique_id128, void** handle);
void mh_comm_destroy(void* handle);
/* attach a communicator to a context of the same rank / world / device.  prev_rank / next_rank: ranks owning the frames before /
 * after this rank's range, -1 at the ends of the sequence */
int mh_set_comm(mh_ctx* ctx, void* handle, int32_t prev_rank, int32_t next_rank);
int mh_has_comm(mh_ctx* ctx);
/* one fit() cycle (optimizer.py:375-587) enqueued on `stream`: [halo exchange] -> mh_fit_grads -> [all-reduce of the shared-leaf
 * gradients and losses] -> mh_fit_update(lr).  A context with world > 1 needs mh_set_comm first. */
int mh_fit_cycle(mh_ctx* ctx, float lr, void* stream);
/* the same without the step (the caller rebuilds the scene between the gradients and mh_fit_update, optimizer.py:578-587) */
int mh_fit_cycle_grads(mh_ctx* ctx, void* stream);
/* one iteration of the translation init (optimizer.py:743-761): [halo] -> mh_init_grads -> [all-reduce] -> mh_init_update */
int mh_init_cycle(mh_ctx* ctx, float lr, int32_t step, void* stream);

/* Device-resident scene update of one cycle >= 30 (optimizer.py:578-584) after the median passes of the depth: median depth ->
 * postprocess_depthmap (utils.py:174-209: bilateral filter, Sobel edge mask, two erosions, fill-in sweeps of utils.py:91-135) ->
 * update_scene_pointcloud (optimizer.py:605-616), all on the device.  depth_host_or_null: (H,W) post-processed depth map. */
int mh_scene_update_from_median(mh_ctx* ctx, int32_t use_bilateral_filter, int32_t fillin_ksize, float* depth_host_or_null, void* stream);
/* postprocess_depthmap (utils.py:174-209) of a host depth map (H,W) on the device; mask_host_or_null (H,W) u8 {0,1} */
int mh_postprocess_depthmap(mh_ctx* ctx, const float* depth_host, const uint8_t* mask_host_or_null, int32_t use_bilateral_filter,
                            int32_t fillin_ksize, float* out_host, void* stream);
int mh_scene_median_finish(mh_ctx* ctx, int32_t which, float* depth_host, uint8_t* mask_host, uint8_t* img_host, void* stream);

/* ---- debugging / input synthesis ---------------------------------------------------------------- */
/* renders person n of local frame t with the current parameters: zbuf[...,0] (depth raster, optimizer.py:429-431)
 * and the soft silhouette alpha (optimizer.py:447-448) as dense (H,W) planes; HOST pointers, blocking. */
int mh_debug_render(mh_ctx* ctx, int32_t t_local, int32_t n, float* zbuf_host, float* alpha_host);
/* fills the context's own depth / seg planes from the CURRENT parameters (ground-truth motion): hard z-buffer
 * per person, nearest person wins, normalised disparity of scene (ground y, wall z) U persons.  Bench / test
 * input synthesis on the device (oracle.synth.assemble_inputs semantics). */
int mh_synth_planes(mh_ctx* ctx, float y_ground, float z_wall, void* stream);
/* copies the context's planes of frames [t_local0, t_local0+count) back to HOST buffers (either may be NULL) */
int mh_read_planes(mh_ctx* ctx, int32_t t_local0, int32_t count, float* depths_host, float* seg_host);
/* CUDA-event timing of the stages of mh_fit_grads on the caller's stream, ring of the last 64 cycles.
 * mh_read_timing: out (n_cycles, 6) ms [smpl forward | pre-raster terms | order prepass | render | smpl backward | post terms],
 * oldest first; *n_cycles is the capacity in and the number of cycles written out.  Blocking. */
int mh_set_timing(mh_ctx* ctx, int32_t on);
int mh_read_timing(mh_ctx* ctx, float* out_ms, int32_t* n_cycles);
/* development aid: counters of the render kernel summed over its CTAs, 32 int64: [0..7] cycles per phase (load+NDC | binning |
 * tile set-up + descriptors | prune + evaluate | per-pixel + silhouette backward | sums + depth backward | chain rule | unused),
 * [8..19] (face, pixel)-pair statistics when the library was built with -DMH_RSTATS (else 0), rest unused;
 * `on` (re)starts / stops counting, out32_host (may be NULL) receives the counters accumulated so far. */
int mh_render_profile(mh_ctx* ctx, int32_t on, long long* out32_host);
/* testing aid: REDUCE the render capacities (0 keeps a value) so that small inputs reach the coarse-binning path (maxbins)
 * and the MH_E_CAPACITY paths (tile-list entries per body, depth-winner entries per body); clears the capacity flag */
int mh_debug_set_render_caps(mh_ctx* ctx, int32_t maxbins, int32_t bincap, int32_t wcap);
/* contact term (optimizer.py:492-500): the 32 nearest scene points of every lowest vertex are searched on a uniform grid over the cloud
 * (rebuilt by mh_set_scene / mh_set_scene_from_depth), exactly, with the streaming search over the whole cloud as the fallback.
 * out[0] = person-frames the grid answered in the last cycle, out[1] = grid cells, out[2] = points, out[3] = cell size (micrometres).
 * The environment variable MH_KNN_GRID=0 (read when the cloud is set) disables the grid. */
int mh_debug_knn_stats(mh_ctx* ctx, int64_t* out4, void* stream);
/* Device buffers of 256 KB and more are recycled between the contexts of a process (a job that fits many sequences, predict.py:315-357,
 * creates one optimiser per sequence): mh_destroy parks them, mh_create takes buffers of the same size back.  mh_pool_trim returns the
 * parked buffers to the driver (done automatically when an allocation fails), mh_pool_bytes says how much is parked (at most 32 GB,
 * MH_POOL_MAX_GB in the environment).  MH_POOL=0 turns the recycling off. */
/* testing aid, host only: float32 {0, 1} masks (count, N, HW) -> one 32-bit plane per frame (bit n = person n), as mh_ingest_frames packs
 * them with all cores; returns 1 when a value other than 0 / 1 was seen (mh_finalize_ingest then fails), 0 otherwise */
int mh_debug_pack_masks(const float* seg_host, int32_t count, int32_t N, int64_t HW, uint32_t* out_host);
void mh_pool_trim(void);
int64_t mh_pool_bytes(void);
/* testing aid: the pose-corrective contraction alone, C (M, 20672) = A (M, 192) . posedirs, HOST pointers, blocking; use_tc = 1 runs
 * the tcgen05 / TMEM kernel (3 x TF32), 0 the FP32 SIMT kernel (smpl.py:549-553 without the shape term) */
int mh_debug_gemm_fwd(mh_ctx* ctx, const float* A_host, float* C_host, int32_t M, int32_t use_tc);
/* testing aid: the backward contraction alone, D (M, 208) = E (M, 20672) . extended_basis^T (dL/dpose_feature and the shape-blend
 * part of dL/dbeta), split-K partials summed on the host; HOST pointers, blocking */
int mh_debug_gemm_bwd(mh_ctx* ctx, const float* E_host, float* D_host, int32_t M, int32_t use_tc);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t mh_launch_count(const mh_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MHOPT_H_ */
