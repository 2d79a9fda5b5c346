// Differentiable mesh rasterisation of every local person-frame, fused with the depth and silhouette
// terms and their analytic backward.
//
// Reference path replaced (mhmocap/optimizer.py): the two PyTorch3D rasterisations per batch -- depth
// (faces_per_pixel 8, blur 1e-4, :211-218, 428-431) and soft silhouette (faces_per_pixel 4, blur 2e-5,
// SoftSilhouetteShader, :221-232, 447-448) -- the erosion / validity masks (:432-438), the average
// log-disparity loss (:440-442, losses.py:19-30), the occlusion-ordered masked-MSE silhouette loss with its
// per-person-frame host sync (:450-475, losses.py:33-40) and PyTorch3D's atomic-add raster backward.
// PyTorch3D semantics: SURVEY.md Appendix A / oracle/raster.py (parity unpinned upstream).
//
// One persistent CTA (1024 threads) per SM pulls person-frames from an atomic counter, largest first.  Per body:
//   P0  the body's 6890 absolute vertices are staged into shared memory with one TMA bulk copy
//       (cp.async.bulk + mbarrier) and converted in place to NDC;
//   P1  faces are binned to 32x32-pixel tiles anchored at the body's bounding box; the tile lists (face ids in a
//       per-CTA global scratch) are counting-sorted by (tile, depth slab) in one fill pass, near -> far;
//   per tile:
//   T0  every thread owns one pixel: the instance bit planes decide whether the loss needs a depth fragment / silhouette
//       fragments there (keys of pixels that need nothing start at 0 and prune every face); one thread per face of the
//       tile builds the (face, tile) descriptor in shared memory (exact pixel rectangle, reciprocals, depth bound);
//   P2  PRUNE: a warp walks the pixel rectangle of one face, 32 pixels per pass, and appends the pixels where the face's
//       nearest vertex is not behind the keys it could displace to a per-warp queue; EVALUATE: every 32 queued survivors
//       -- from whichever faces -- are evaluated with all lanes busy (depth first, distance test only if a key would be
//       displaced) and update the per-pixel keys (depth | face): one 64-bit min for the nearest depth fragment, a 4-slot
//       concurrent sorted insertion for the four nearest silhouette fragments; both rasters share one evaluation;
//   P3  per pixel: values of the <= 5 winning fragments from the descriptors, depth-loss sums, silhouette alpha, loss and
//       its backward (fire-and-forget reductions into a per-CTA gradient row that stays in L2); depth winners go to a
//       compact list because their gradient scale needs the whole-image sums;
//   P4  block reduction of the sums, depth backward over the winner list;
//   P5  NDC gradients are chained to camera space and added to dL/dV.
// Everything outside the tiles a body touches contributes a mesh-independent constant that the prepass
// (mh_planes.cu) keeps per (frame, order position).
#include "mh_ctx.h"

#define TW 32                 // tile width  (pixels)
#define TH 32                 // tile height (pixels)
#define R_THREADS 1024        // one thread per tile pixel in the per-pixel phases
#define R_MAXBINS 1024
#define R_NSLAB 64            // depth slabs (at most): the tile lists are ordered near -> far so that later faces are pruned early
#define KEY_EMPTY 0xffffffffffffffffull
#define R_DESC 1024               // (face, tile) descriptors staged per chunk (5 float4 each)
#define R_QUEUE 64                // survivor queue entries per warp (fewer than 32 wait, a pass adds at most 32)
#ifndef MH_R_GRAB
#define MH_R_GRAB 1               // faces a warp takes from the tile's list per hand-out (one shared-memory atomic); measured 1 < 2 < 3 < 4 (balance at the tile barrier)
#endif

struct MhRenderScratch {
    int* cost; int* work; bool have_cost;        // cycles each body took in the previous launch, bodies sorted by them (longest first)
    uint16_t* binlist; int bincap;
    uint2* fbin;
    int* wpix; int* wface; float* wz; int wcap;
    int nctas;
    size_t smem;
    int* counter;
    long long* gsg; int nslab;
    long long* prof;
    int maxbins;               // tile bins per body before the binning granularity is coarsened (<= R_MAXBINS)
    int bincap_use, wcap_use;  // capacities handed to the kernel (<= the allocated ones; testing aid)
};

struct RenderParams {
    const float* verts; float* dverts; const int32_t* faces;
    const float* pix_x; const float* pix_y;
    const float* depth; const uint32_t* cbits; const uint32_t* ebits;
    const int* order; const uint32_t* premask; const int* rankcnt;
    const uint8_t* pose2d_valid; const uint8_t* mask_valid;
    const float* zmin_lin; const float* zmax_lin;
    float* pfout; int* devflags;
    int* cost;                    // out: clock cycles this launch spent on body i (feeds the next launch's order)
    const int* work;              // bodies in the order they are handed out (longest first by the previous launch's cycles); NULL: by depth rank
    int nslab;                    // depth slabs of the tile lists (<= 256)
    long long* gsg;               // per-CTA NDC-gradient rows (MH_LD3V 64-bit fixed-point sums each), zero between bodies
    uint16_t* binlist; int bincap;
    uint2* fbin;                  // per-CTA scratch: packed bin range + depth slab per face
    int* wpix; int* wface; float* wz; int wcap;
    int* counter;
    int T, N, H, W;
    float k00, k02, k11, k12;
    float rx, ry;                 // NDC extent of the x / y axis (2 on the short side)
    float blur_d, blur_s, r_d, r_s, sigma, eps;    // r_s: conservative (x 1.001) radius of the silhouette raster
    float coef_depth, coef_sil;
    float* dbg_zbuf; float* dbg_alpha; int dbg_body;
    int maxbins;
    long long* prof;              // optional: per-phase cycle counters (8 per CTA)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float wsum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// fractional pixel index of an NDC coordinate on an axis of S pixels and NDC extent r (inverse of the
// pixel-centre formula; used only for conservative ranges)
__device__ __forceinline__ float pix_of(float ndc, int S, float r) { return (float)(S - 1) - ((ndc + 0.5f * r) * (float)S - 0.5f * r) / r; }

// approximate squared distance to a segment (reciprocal multiply instead of the oracle's division); only used
// to decide far from the blur threshold -- near it the exact form is evaluated
__device__ __forceinline__ float seg_dist_fast(float dx, float dy, float bax, float bay, float il, float dbx, float dby) {
    if (il == 0.f) return dbx * dbx + dby * dby;
    const float t = __saturatef((bax * dx + bay * dy) * il);
    const float qx = dx - t * bax, qy = dy - t * bay;
    return qx * qx + qy * qy;
}

__device__ __forceinline__ void load_face(const float* sv, const int32_t* __restrict__ faces, int f, float r, MhFace* fc, int iv[3]) {
    iv[0] = faces[3 * f]; iv[1] = faces[3 * f + 1]; iv[2] = faces[3 * f + 2];
    mh_face_setup(sv + 3 * iv[0], sv + 3 * iv[1], sv + 3 * iv[2], r, fc);
}

// Shared-memory accesses of the hot loops, in the shared state space with 32-bit window addresses.  Through generic pointers
// the compiler rebuilds the generic base of the dynamic array (S2R SR_CgaCtaId, shifts, adds) next to the accesses, inside the
// pair loop; the atomics additionally expand to match / elect / popc sequences.
template <int OFF> __device__ __forceinline__ unsigned lds32(uint32_t a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ float ldsf(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ unsigned long long lds64(uint32_t a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1+%2];" : "=l"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ void sts32(uint32_t a, unsigned v) { asm volatile("st.shared.u32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v) : "memory"); }
template <int OFF> __device__ __forceinline__ unsigned long long atoms_min64(uint32_t a, unsigned long long v) {
    unsigned long long old;
    asm volatile("atom.shared.min.u64 %0, [%1+%2], %3;" : "=l"(old) : "r"(a), "n"(OFF), "l"(v) : "memory");
    return old;
}
__device__ __forceinline__ int atoms_inc(uint32_t a) {        // old value, then + 1 (ptxas leaves .inc alone; a uniform-address .add becomes vote + popc + shuffle)
    int old;
    asm volatile("atom.shared.inc.u32 %0, [%1], 0xffffffff;" : "=r"(old) : "r"(a) : "memory");
    return old;
}
__device__ __forceinline__ int atoms_add(uint32_t a, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}

// offsets of the dynamic shared-memory regions (bytes)
constexpr int SO_DKEY = MH_LD3V * 4;                                   // R_THREADS     nearest depth fragment
constexpr int SO_SKEY = SO_DKEY + R_THREADS * 8;                       // 4 x R_THREADS nearest silhouette fragments
constexpr int SO_TCOUNT = SO_SKEY + 4 * R_THREADS * 8;                 // R_MAXBINS + 1 (exclusive offsets after the scan)
constexpr int SO_TCUR = SO_TCOUNT + (R_MAXBINS + 1) * 4;               // R_MAXBINS (+ 3 pad)
constexpr int SO_SRED = SO_TCUR + (R_MAXBINS + 3) * 4;                 // 256 floats
constexpr int SO_SPX = SO_SRED + 256 * 4;                              // TW
constexpr int SO_SPY = SO_SPX + TW * 4;                                // TH
constexpr int SO_SINT = SO_SPY + TH * 4;                               // 64 ints
constexpr int SO_SDESC = (SO_SINT + 64 * 4 + 15) & ~15;                // R_DESC x 5 float4
constexpr int SO_QUEUE = SO_SDESC + R_DESC * 80;                       // per-warp survivor queues (R_QUEUE words each)
constexpr int SO_END = SO_QUEUE + (R_THREADS / 32) * R_QUEUE * 4;
constexpr int SK_STRIDE = R_THREADS * 8;                               // bytes between the slot planes of the keys

__device__ __forceinline__ void key_insert4(uint32_t a /* slot 0 of the pixel */, unsigned long long x) {
    // concurrent sorted insertion: every slot keeps the minimum of what reaches it and passes the rest on
    unsigned long long old = atoms_min64<0>(a, x);
    x = old > x ? old : x;
    if (x == KEY_EMPTY) return;
    old = atoms_min64<SK_STRIDE>(a, x);
    x = old > x ? old : x;
    if (x == KEY_EMPTY) return;
    old = atoms_min64<2 * SK_STRIDE>(a, x);
    x = old > x ? old : x;
    if (x == KEY_EMPTY) return;
    atoms_min64<3 * SK_STRIDE>(a, x);
}

// P3 on single-chunk tiles: the same fragment values / silhouette backward from the tile's (face, tile) descriptor that is still in
// shared memory -- no index or vertex gathers, no reciprocals (the descriptor holds the ones P2 used)
__device__ __forceinline__ void frag_values_desc(uint32_t da, float px, float py, float* pz, float* sd) {
    const float4 q0 = lds128<0>(da), q1 = lds128<16>(da), q2 = lds128<32>(da);
    const float x0 = q0.x, y0 = q0.y, x1 = q0.z, y1 = q0.w, x2 = q1.x, y2 = q1.y, z0 = q1.z, z1 = q1.w, z2 = q2.x;
    const float inv_den = q2.y;
    const float dx0 = px - x0, dy0 = py - y0, dx1 = px - x1, dy1 = py - y1, dx2 = px - x2, dy2 = py - y2;
    const float ex12 = x2 - x1, ey12 = y2 - y1, ex20 = x0 - x2, ey20 = y0 - y2, ex01 = x1 - x0, ey01 = y1 - y0;
    const float e0 = MH_SUB(MH_MUL(dx1, ey12), MH_MUL(dy1, ex12));
    const float e1 = MH_SUB(MH_MUL(dx2, ey20), MH_MUL(dy2, ex20));
    const float e2 = MH_SUB(MH_MUL(dx0, ey01), MH_MUL(dy0, ex01));
    const bool dpos = inv_den > 0.f;
    const bool inside = (e0 != 0.f) && (e1 != 0.f) && (e2 != 0.f) && ((e0 > 0.f) == dpos) && ((e1 > 0.f) == dpos) && ((e2 > 0.f) == dpos);
    const float c0 = __saturatef(e0 * inv_den), c1 = __saturatef(e1 * inv_den), c2 = __saturatef(e2 * inv_den);
    *pz = (c0 * z0 + c1 * z1 + c2 * z2) / fmaxf(c0 + c1 + c2, 1e-5f);
    const float d01 = seg_dist_fast(dx0, dy0, ex01, ey01, q2.z, dx1, dy1);
    const float d02 = seg_dist_fast(dx0, dy0, -ex20, -ey20, q2.w, dx2, dy2);
    const float d12 = seg_dist_fast(dx1, dy1, ex12, ey12, ldsf<64>(da), dx2, dy2);
    const float d = fminf(fminf(d01, d02), d12);
    *sd = inside ? -d : d;
}

__device__ __forceinline__ void grad_add(long long* p, float v);

__device__ __forceinline__ void sil_grad_desc(long long* sg, uint32_t da, float px, float py, float gd) {
    const float4 q0 = lds128<0>(da), q1 = lds128<16>(da), q2 = lds128<32>(da), q4 = lds128<64>(da);
    const float vx[3] = {q0.x, q0.z, q1.x}, vy[3] = {q0.y, q0.w, q1.y};
    const float il[3] = {q2.z, q2.w, q4.x};                              // edges 01, 02, 12
    const unsigned i01 = __float_as_uint(q4.z);
    const int iv[3] = {(int)(i01 & 0xffffu), (int)(i01 >> 16), __float_as_int(q4.w)};
    float best = INFINITY, bt = 0.f, bqx = 0.f, bqy = 0.f;
    int ia = 0, ib = 1;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int a = (e == 2) ? 1 : 0, b = (e == 0) ? 1 : 2;
        const float bax = vx[b] - vx[a], bay = vy[b] - vy[a];
        const float t = (il[e] == 0.f) ? 1.0f : __saturatef((bax * (px - vx[a]) + bay * (py - vy[a])) * il[e]);
        const float qx = px - (vx[a] + t * bax), qy = py - (vy[a] + t * bay);
        const float d = qx * qx + qy * qy;
        if (d < best) { best = d; bt = t; bqx = qx; bqy = qy; ia = iv[a]; ib = iv[b]; }
    }
    const float g = -2.0f * gd;
    const float ga = g * (1.0f - bt), gb = g * bt;
    if (ga != 0.f) { grad_add(&sg[3 * ia], ga * bqx); grad_add(&sg[3 * ia + 1], ga * bqy); }
    if (gb != 0.f) { grad_add(&sg[3 * ib], gb * bqx); grad_add(&sg[3 * ib + 1], gb * bqy); }
}

// -DMH_RSTATS: (face, pixel)-pair statistics of P2 in slots 8.. of the profile buffer (instrumented build only)
#ifdef MH_RSTATS
#define RS_DECL long long rs_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define RS_ADD(k, v) (rs_[k] += (v))
#define RS_WARP(k) do { const unsigned am_ = __activemask(); if ((am_ & (0u - am_)) == (1u << lane)) rs_[k] += 1; } while (0)
#define RS_FLUSH do { if (P.prof) for (int k_ = 0; k_ < 12; ++k_) if (rs_[k_]) atomicAdd((unsigned long long*)&P.prof[blockIdx.x * MH_NPROF + 8 + k_], (unsigned long long)rs_[k_]); } while (0)
#else
#define RS_DECL
#define RS_ADD(k, v)
#define RS_WARP(k)
#define RS_FLUSH
#endif
#define MH_NPROF 32

#define PROF(k) do { if (P.prof && tid == 0) { const long long now_ = clock64(); P.prof[blockIdx.x * MH_NPROF + (k)] += now_ - tprof; tprof = now_; } } while (0)

// depth and signed squared edge distance of one (face, pixel) fragment with reciprocal multiplies (values only: every
// DECISION was taken in P2 with the oracle's exact arithmetic)
__device__ __forceinline__ void frag_values(const float* sv, const int32_t* __restrict__ faces, int f, float px, float py, float* pz, float* sd) {
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float x0 = sv[3 * i0], y0 = sv[3 * i0 + 1], z0 = sv[3 * i0 + 2];
    const float x1 = sv[3 * i1], y1 = sv[3 * i1 + 1], z1 = sv[3 * i1 + 2];
    const float x2 = sv[3 * i2], y2 = sv[3 * i2 + 1], z2 = sv[3 * i2 + 2];
    const float den = MH_ADD(mh_edge(x2, y2, x0, y0, x1, y1), MH_KEPS);
    const float inv_den = __frcp_rn(den);
    const float dx0 = px - x0, dy0 = py - y0, dx1 = px - x1, dy1 = py - y1, dx2 = px - x2, dy2 = py - y2;
    const float ex12 = x2 - x1, ey12 = y2 - y1, ex20 = x0 - x2, ey20 = y0 - y2, ex01 = x1 - x0, ey01 = y1 - y0;
    const float e0 = MH_SUB(MH_MUL(dx1, ey12), MH_MUL(dy1, ex12));
    const float e1 = MH_SUB(MH_MUL(dx2, ey20), MH_MUL(dy2, ex20));
    const float e2 = MH_SUB(MH_MUL(dx0, ey01), MH_MUL(dy0, ex01));
    const bool dpos = den > 0.f;
    const bool inside = (e0 != 0.f) && (e1 != 0.f) && (e2 != 0.f) && ((e0 > 0.f) == dpos) && ((e1 > 0.f) == dpos) && ((e2 > 0.f) == dpos);
    const float c0 = __saturatef(e0 * inv_den), c1 = __saturatef(e1 * inv_den), c2 = __saturatef(e2 * inv_den);
    *pz = (c0 * z0 + c1 * z1 + c2 * z2) / fmaxf(c0 + c1 + c2, 1e-5f);
    const float l01 = ex01 * ex01 + ey01 * ey01, l02 = ex20 * ex20 + ey20 * ey20, l12 = ex12 * ex12 + ey12 * ey12;
    const float d01 = seg_dist_fast(dx0, dy0, ex01, ey01, l01 <= MH_KEPS ? 0.f : __frcp_rn(l01), dx1, dy1);
    const float d02 = seg_dist_fast(dx0, dy0, -ex20, -ey20, l02 <= MH_KEPS ? 0.f : __frcp_rn(l02), dx2, dy2);
    const float d12 = seg_dist_fast(dx1, dy1, ex12, ey12, l12 <= MH_KEPS ? 0.f : __frcp_rn(l12), dx2, dy2);
    const float d = fminf(fminf(d01, d02), d12);
    *sd = inside ? -d : d;
}

// gradient scatter: shared-memory float atomics are compare-and-swap loops (ATOMS.CAST.SPIN) that retry when the pixels of a
// warp hit the same vertex; the gradients go instead as fire-and-forget reductions to a per-CTA row that stays in L2.  The sums are
// 64-bit FIXED-POINT (2^-44 NDC-gradient units, +-5e5 range): integer addition is associative, so the per-vertex sum does not depend
// on the order in which the pixels of the body arrive -- float reductions made every cycle differ in the last bits from run to run
#define MH_GFIX 17592186044416.0f          // 2^44
__device__ __forceinline__ void grad_add(long long* p, float v) {
    const long long q = __float2ll_rn(fminf(fmaxf(v, -131072.0f), 131072.0f) * MH_GFIX);
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(q) : "memory");
}


// Descriptor of one (face, tile) item -- everything P2 needs, computed ONCE by one thread (the tile's faces are spread over
// the 1024 threads) instead of redundantly by the 32 lanes of the warp that rasterises the face:
//   d0 = x0 y0 x1 y1 | d1 = x2 y2 z0 z1 | d2 = z2 1/den 1/|e01|^2 1/|e02|^2 | d3 = span inner zbits lanes | d4 = 1/|e12|^2 face i0|i1<<16 i2
// d3    = span inner zbits lanes (layouts in make_desc): the face's pixel rectangle inside the tile, EXACT for the oracle's bbox test
//         (bbox inflated by sqrt(blur) of the depth raster), so the pair loop needs no per-pixel bbox test, and the INNER rectangle,
//         the only pixels where the face can be a SILHOUETTE fragment (bbox inflated by the silhouette radius x 1.001 and 0.01 px:
//         conservative, the exact distance test follows per pixel)
// zbits = bits of a lower bound of every fragment depth of the face (its nearest vertex)
__device__ __forceinline__ void make_desc(const RenderParams& P, const float* sv, int f, int ox, int oy, int txmax, int tymax, float4* d) {
    const int i0 = P.faces[3 * f], i1 = P.faces[3 * f + 1], i2 = P.faces[3 * f + 2];
    const float x0 = sv[3 * i0], y0 = sv[3 * i0 + 1], z0 = sv[3 * i0 + 2];
    const float x1 = sv[3 * i1], y1 = sv[3 * i1 + 1], z1 = sv[3 * i1 + 2];
    const float x2 = sv[3 * i2], y2 = sv[3 * i2 + 1], z2 = sv[3 * i2 + 2];
    const float bxmin = MH_SUB(fminf(fminf(x0, x1), x2), P.r_d), bxmax = MH_ADD(fmaxf(fmaxf(x0, x1), x2), P.r_d);
    const float bymin = MH_SUB(fminf(fminf(y0, y1), y2), P.r_d), bymax = MH_ADD(fmaxf(fmaxf(y0, y1), y2), P.r_d);
    // pixel rectangle of the inflated bbox (conservative by 0.01 px), clipped to the tile, then made exact
    int c0 = max((int)fmaxf(ceilf(pix_of(bxmax, P.W, P.rx) - 0.01f), 0.f), ox), c1 = min((int)fminf(floorf(pix_of(bxmin, P.W, P.rx) + 0.01f), (float)(P.W - 1)), ox + txmax);
    int r0 = max((int)fmaxf(ceilf(pix_of(bymax, P.H, P.ry) - 0.01f), 0.f), oy), r1 = min((int)fminf(floorf(pix_of(bymin, P.H, P.ry) + 0.01f), (float)(P.H - 1)), oy + tymax);
    while (c0 <= c1 && (P.pix_x[c0] > bxmax || P.pix_x[c0] < bxmin)) ++c0;
    while (c1 >= c0 && (P.pix_x[c1] > bxmax || P.pix_x[c1] < bxmin)) --c1;
    while (r0 <= r1 && (P.pix_y[r0] > bymax || P.pix_y[r0] < bymin)) ++r0;
    while (r1 >= r0 && (P.pix_y[r1] > bymax || P.pix_y[r1] < bymin)) --r1;
    if (c0 > c1 || r0 > r1) { d[3] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    int ic0 = max((int)fmaxf(ceilf(pix_of(bxmax - P.r_d + P.r_s, P.W, P.rx) - 0.01f), 0.f), c0), ic1 = min((int)fminf(floorf(pix_of(bxmin + P.r_d - P.r_s, P.W, P.rx) + 0.01f), (float)(P.W - 1)), c1);
    int ir0 = max((int)fmaxf(ceilf(pix_of(bymax - P.r_d + P.r_s, P.H, P.ry) - 0.01f), 0.f), r0), ir1 = min((int)fminf(floorf(pix_of(bymin + P.r_d - P.r_s, P.H, P.ry) + 0.01f), (float)(P.H - 1)), r1);
    // the three words the prune passes decode (pass = whole rows of the rectangle, 32 / w of them; lane -> (row lr, column) once per face):
    //   span  = first pixel r0 * TW + c0 | end (r0 + h) * TW << 10 | pixels per pass rpp * TW << 21          (0: nothing in this tile)
    //   inner = first inner pixel row jr0 * TW | inner rows jh * TW << 10 | first inner column - c0 << 21 | inner columns << 26
    //   lanes = ceil(65536 / w) (lane / w in 16.16 fixed point) | TW - w << 17 | lanes in use rpp * w << 23
    const int w = c1 - c0 + 1, h = r1 - r0 + 1, rpp = 32 / w;
    const unsigned span = (unsigned)((r0 - oy) * TW + (c0 - ox)) | ((unsigned)((r0 - oy + h) * TW) << 10) | ((unsigned)(rpp * TW) << 21);
    unsigned inner = 0u;
    if (ic0 <= ic1 && ir0 <= ir1) inner = (unsigned)((ir0 - oy) * TW) | ((unsigned)((ir1 - ir0 + 1) * TW) << 10) | ((unsigned)(ic0 - c0) << 21) | ((unsigned)(ic1 - ic0 + 1) << 26);
    const unsigned lanes = (unsigned)((65536 + w - 1) / w) | ((unsigned)(TW - w) << 17) | ((unsigned)(rpp * w) << 23);
    const float den = MH_ADD(mh_edge(x2, y2, x0, y0, x1, y1), MH_KEPS);
    const float ex12 = x2 - x1, ey12 = y2 - y1, ex20 = x0 - x2, ey20 = y0 - y2, ex01 = x1 - x0, ey01 = y1 - y0;
    const float l01 = ex01 * ex01 + ey01 * ey01, l02 = ex20 * ex20 + ey20 * ey20, l12 = ex12 * ex12 + ey12 * ey12;
    const float zmin = fminf(z0, fminf(z1, z2));
    d[0] = make_float4(x0, y0, x1, y1);
    d[1] = make_float4(x2, y2, z0, z1);
    d[2] = make_float4(z2, __frcp_rn(den), l01 <= MH_KEPS ? 0.f : __frcp_rn(l01), l02 <= MH_KEPS ? 0.f : __frcp_rn(l02));
    d[3] = make_float4(__uint_as_float(span), __uint_as_float(inner), __uint_as_float(__float_as_uint(fmaxf(zmin * (1.0f - 1e-6f), 0.f))), __uint_as_float(lanes));
    d[4] = make_float4(l12 <= MH_KEPS ? 0.f : __frcp_rn(l12), __int_as_float(f), __uint_as_float((unsigned)i0 | ((unsigned)i1 << 16)), __int_as_float(i2));
}

// Backward of the unsigned squared edge distance of one silhouette fragment: d = |p - a - t (b - a)|^2 on the nearest edge
// (first minimum in the order 01, 02, 12, as mh_face_bwd), dd/da = -2 q (1 - t), dd/db = -2 q t.  Reciprocal multiplies: the
// gradient tolerance (1e-3 of the maximum) does not need the oracle's divisions.
__device__ __forceinline__ void sil_grad(long long* sg, const float* sv, const int32_t* __restrict__ faces, int f, float px, float py, float gd) {
    const int iv[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
    const float vx[3] = {sv[3 * iv[0]], sv[3 * iv[1]], sv[3 * iv[2]]};
    const float vy[3] = {sv[3 * iv[0] + 1], sv[3 * iv[1] + 1], sv[3 * iv[2] + 1]};
    float best = INFINITY, bt = 0.f, bqx = 0.f, bqy = 0.f;
    int ia = 0, ib = 1;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int a = (e == 2) ? 1 : 0, b = (e == 0) ? 1 : 2;
        const float bax = vx[b] - vx[a], bay = vy[b] - vy[a];
        const float l2 = bax * bax + bay * bay;
        const float t = (l2 <= MH_KEPS) ? 1.0f : __saturatef(__fdividef(bax * (px - vx[a]) + bay * (py - vy[a]), l2));
        const float qx = px - (vx[a] + t * bax), qy = py - (vy[a] + t * bay);
        const float d = qx * qx + qy * qy;
        if (d < best) { best = d; bt = t; bqx = qx; bqy = qy; ia = a; ib = b; }
    }
    const float g = -2.0f * gd;
    const float ga = g * (1.0f - bt), gb = g * bt;
    if (ga != 0.f) { grad_add(&sg[3 * iv[ia]], ga * bqx); grad_add(&sg[3 * iv[ia] + 1], ga * bqy); }
    if (gb != 0.f) { grad_add(&sg[3 * iv[ib]], gb * bqx); grad_add(&sg[3 * iv[ib] + 1], gb * bqy); }
}

template <int MODE>      // MODE 0: losses + gradients ; 1: dense zbuf / alpha planes of one body
__global__ void __launch_bounds__(R_THREADS, 1) k_render(RenderParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sv = reinterpret_cast<float*>(smem_raw);                                   // MH_LD3V  NDC vertices
    long long* sg = P.gsg + (size_t)blockIdx.x * MH_LD3V;                             // MH_LD3V  NDC gradients (fixed point): global (L2), zero on entry
    unsigned long long* dkey = reinterpret_cast<unsigned long long*>(smem_raw + SO_DKEY);
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem_raw + SO_SKEY);
    int* tcount = reinterpret_cast<int*>(smem_raw + SO_TCOUNT);
    float* sred = reinterpret_cast<float*>(smem_raw + SO_SRED);
    float* spx = reinterpret_cast<float*>(smem_raw + SO_SPX);
    float* spy = reinterpret_cast<float*>(smem_raw + SO_SPY);
    int* sint = reinterpret_cast<int*>(smem_raw + SO_SINT);
    float4* sdesc = reinterpret_cast<float4*>(smem_raw + SO_SDESC);
    // shared-window address of the dynamic array.  ptxas builds it from SR_CgaCtaId (S2R + 3 ALU) and would rematerialise that
    // inside the pair loop; a volatile asm pins it to ONE evaluation
    uint32_t sb;
    asm volatile("mov.u32 %0, smem_raw;" : "=r"(sb));
    __shared__ __align__(8) unsigned long long mbar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = R_THREADS / 32;
    const int TN = P.T * P.N;
    uint16_t* binlist = P.binlist + (size_t)blockIdx.x * P.bincap;
    uint2* fbin = P.fbin + (size_t)blockIdx.x * MH_F;
    int* wpix = P.wpix + (size_t)blockIdx.x * P.wcap;
    int* wface = P.wface + (size_t)blockIdx.x * P.wcap;
    float* wz = P.wz + (size_t)blockIdx.x * P.wcap;
    const float blur_d_lo = P.blur_d * (1.0f - 1e-5f), blur_d_hi = P.blur_d * (1.0f + 1e-5f);
    const float blur_s_lo = P.blur_s * (1.0f - 1e-5f), blur_s_hi = P.blur_s * (1.0f + 1e-5f);
    uint32_t phase = 0;
    long long tprof = clock64();
    RS_DECL;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    for (int iter = 0;; ++iter) {
        // ---- next body (dynamic scheduling: bodies differ a lot in projected size) ----
        if (MODE == 1 && iter > 0) break;
        if (tid == 0) sint[0] = (MODE == 1) ? P.dbg_body : atomicAdd(P.counter, 1);
        __syncthreads();
        const int wi = sint[0];
        __syncthreads();
        if (wi >= TN) break;
        // longest-first: work item w = (depth-order position q, frame t) -> the nearest (largest) persons of all frames are
        // rasterised first, the small far ones fill the tail
        int t, n;
        if (MODE == 1) { t = wi / P.N; n = wi % P.N; }
        else if (P.work) { const int iw = P.work[wi]; t = iw / P.N; n = iw - t * P.N; }
        else { t = wi % P.T; n = P.order[t * P.N + wi / P.T]; }
        const int i = t * P.N + n;
        if (MODE == 0 && tid == 0) { const long long now = clock64(); sint[42] = (int)(now & 0xffffffffll); sint[43] = (int)(now >> 32); }
        const size_t b = (size_t)i + P.N;                                // slot-major body (slot 0 is the halo)
        // ---- P0: TMA bulk copy of the vertex row, then world -> NDC in place ----
        if (tid == 0) {
            const uint32_t bytes = MH_LD3V * sizeof(float);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sv)),
                         "l"(P.verts + b * MH_LD3V), "r"(bytes), "r"(smem_u32(&mbar))
                         : "memory");
        }
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
            }
        }
        phase ^= 1;
        float bx0 = INFINITY, bx1 = -INFINITY, by0 = INFINITY, by1 = -INFINITY, bz0 = INFINITY, bz1 = -INFINITY;
        for (int v = tid; v < MH_V; v += R_THREADS) {
            const float Pw[3] = {sv[3 * v], sv[3 * v + 1], sv[3 * v + 2]};
            float o[3];
            mh_world_to_ndc(Pw, P.k00, P.k02, P.k11, P.k12, o);
            sv[3 * v] = o[0]; sv[3 * v + 1] = o[1]; sv[3 * v + 2] = o[2];
            if (o[2] > 0.f) { bx0 = fminf(bx0, o[0]); bx1 = fmaxf(bx1, o[0]); by0 = fminf(by0, o[1]); by1 = fmaxf(by1, o[1]); bz0 = fminf(bz0, o[2]); bz1 = fmaxf(bz1, o[2]); }
        }
        for (int o = 16; o > 0; o >>= 1) {
            bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
            bz0 = fminf(bz0, __shfl_xor_sync(0xffffffffu, bz0, o)); bz1 = fmaxf(bz1, __shfl_xor_sync(0xffffffffu, bz1, o));
        }
        if (lane == 0) { sred[warp] = bx0; sred[NW + warp] = bx1; sred[2 * NW + warp] = by0; sred[3 * NW + warp] = by1; sred[4 * NW + warp] = bz0; sred[5 * NW + warp] = bz1; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NW; ++w) {
                bx0 = fminf(bx0, sred[w]); bx1 = fmaxf(bx1, sred[NW + w]); by0 = fminf(by0, sred[2 * NW + w]); by1 = fmaxf(by1, sred[3 * NW + w]);
                bz0 = fminf(bz0, sred[4 * NW + w]); bz1 = fmaxf(bz1, sred[5 * NW + w]);
            }
            // pixel bbox of the body (NDC x / y decrease with the pixel index), inflated by the blur radius + 1 px
            int c0 = (int)floorf(fminf(fmaxf(pix_of(bx1 + P.r_d, P.W, P.rx) - 1.f, 0.f), (float)(P.W - 1)));
            int c1 = (int)ceilf(fminf(fmaxf(pix_of(bx0 - P.r_d, P.W, P.rx) + 1.f, 0.f), (float)(P.W - 1)));
            int r0 = (int)floorf(fminf(fmaxf(pix_of(by1 + P.r_d, P.H, P.ry) - 1.f, 0.f), (float)(P.H - 1)));
            int r1 = (int)ceilf(fminf(fmaxf(pix_of(by0 - P.r_d, P.H, P.ry) + 1.f, 0.f), (float)(P.H - 1)));
            if (!(bx0 <= bx1)) { c0 = 1; c1 = 0; r0 = 1; r1 = 0; }          // nothing in front of the camera
            // the tile grid starts at the body's own bounding box (not at multiples of the tile size): fewer, fuller tiles
            const int tx0 = c0, ty0 = r0;
            int ntx = (c1 >= c0) ? (c1 - c0) / TW + 1 : 0, nty = (r1 >= r0) ? (r1 - r0) / TH + 1 : 0;
            int ks = 0;
            while ((((ntx + (1 << ks) - 1) >> ks) * ((nty + (1 << ks) - 1) >> ks)) > P.maxbins) ++ks;
            // depth slabs per bin: the tile lists are counting-sorted by (bin, slab) in ONE fill pass; as many slabs as the
            // counter array holds
            const int nb = max(((ntx + (1 << ks) - 1) >> ks) * ((nty + (1 << ks) - 1) >> ks), 1);
            const int S = max(min(P.nslab, (2 * R_MAXBINS) / nb), 1);
            sred[6 * NW] = bz0; sred[6 * NW + 1] = (bz1 > bz0) ? (float)S / (bz1 - bz0) : 0.f;
            sint[1] = tx0; sint[2] = ty0; sint[3] = ntx; sint[4] = nty; sint[5] = ks; sint[41] = S;
            sint[6] = 0;     // winner count
            sint[7] = 0;     // overflow flag
        }
        __syncthreads();
        PROF(0);
        const int tx0 = sint[1], ty0 = sint[2], ntx = sint[3], nty = sint[4], ks = sint[5];
        const int nbx = (ntx + (1 << ks) - 1) >> ks, nby = (nty + (1 << ks) - 1) >> ks, nbins = nbx * nby;
        const int S = sint[41], K = nbins * S;                           // counters: (bin, slab), bin-major
        // ---- P1: bin the faces ----
        for (int e = tid; e <= K; e += R_THREADS) tcount[e] = 0;
        __syncthreads();
        const float zlo = sred[6 * NW], zscale = sred[6 * NW + 1];
        // pass 0: bin range (conservative by 0.01 px) + depth slab per face, per-bin counts
        for (int f = tid; f < MH_F; f += R_THREADS) {
            const int i0 = P.faces[3 * f], i1 = P.faces[3 * f + 1], i2 = P.faces[3 * f + 2];
            const float x0 = sv[3 * i0], y0 = sv[3 * i0 + 1], z0 = sv[3 * i0 + 2];
            const float x1 = sv[3 * i1], y1 = sv[3 * i1 + 1], z1 = sv[3 * i1 + 2];
            const float x2 = sv[3 * i2], y2 = sv[3 * i2 + 1], z2 = sv[3 * i2 + 2];
            const float zmax = fmaxf(z0, fmaxf(z1, z2)), zmin = fminf(z0, fminf(z1, z2));
            const float area = mh_edge(x2, y2, x0, y0, x1, y1);
            uint2 fb = make_uint2(0u, 0u);
            if ((zmax >= 0.f) && !((area <= MH_KEPS) && (area >= -MH_KEPS)) && nbins > 0) {
                const float bxmin = MH_SUB(fminf(fminf(x0, x1), x2), P.r_d), bxmax = MH_ADD(fmaxf(fmaxf(x0, x1), x2), P.r_d);
                const float bymin = MH_SUB(fminf(fminf(y0, y1), y2), P.r_d), bymax = MH_ADD(fmaxf(fmaxf(y0, y1), y2), P.r_d);
                const float pc0 = ceilf(pix_of(bxmax, P.W, P.rx) - 0.01f), pc1 = floorf(pix_of(bxmin, P.W, P.rx) + 0.01f);
                const float pr0 = ceilf(pix_of(bymax, P.H, P.ry) - 0.01f), pr1 = floorf(pix_of(bymin, P.H, P.ry) + 0.01f);
                if ((pc1 >= 0.f) && (pr1 >= 0.f) && (pc0 <= (float)(P.W - 1)) && (pr0 <= (float)(P.H - 1)) && (pc0 <= pc1) && (pr0 <= pr1)) {
                    const int c0 = (int)fmaxf(pc0, 0.f), c1 = (int)fminf(pc1, (float)(P.W - 1));
                    const int r0 = (int)fmaxf(pr0, 0.f), r1 = (int)fminf(pr1, (float)(P.H - 1));
                    const int bx_lo = max(((c0 - tx0) / TW) >> ks, 0), bx_hi = min(((c1 - tx0) / TW) >> ks, nbx - 1);
                    const int by_lo = max(((r0 - ty0) / TH) >> ks, 0), by_hi = min(((r1 - ty0) / TH) >> ks, nby - 1);
                    if (bx_lo <= bx_hi && by_lo <= by_hi) {
                        const int slab = min(max((int)((zmin - zlo) * zscale), 0), S - 1);
                        fb = make_uint2((unsigned)bx_lo | ((unsigned)bx_hi << 16), (unsigned)by_lo | ((unsigned)by_hi << 10) | ((unsigned)slab << 20) | 0x80000000u);
                        for (int by = by_lo; by <= by_hi; ++by)
                            for (int bx = bx_lo; bx <= bx_hi; ++bx) atoms_add(sb + SO_TCOUNT + 4 * ((by * nbx + bx) * S + slab), 1);
                    }
                }
            }
            fbin[f] = fb;
        }
        __syncthreads();
        {
            // exclusive scan of tcount[0..K) -> start offsets of every (bin, slab) run
            const int per = (K + R_THREADS - 1) / R_THREADS;
            int local = 0;
            for (int k = 0; k < per; ++k) { const int e = tid * per + k; if (e < K) local += tcount[e]; }
            int incl = local;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            if (lane == 31) sint[8 + warp] = incl;
            __syncthreads();
            int wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += sint[8 + w];
            int run = wbase + incl - local;
            for (int k = 0; k < per; ++k) {
                const int e = tid * per + k;
                if (e < K) { const int cnt = tcount[e]; tcount[e] = run; run += cnt; }
            }
            if (tid == R_THREADS - 1 && run > P.bincap) sint[7] = 1;
            __syncthreads();
        }
        // pass 1: fill -- the start offsets serve as cursors, so afterwards tcount[k] is the END of run k (= start of run k + 1):
        // bin b owns binlist[ (b ? tcount[b S - 1] : 0) .. tcount[b S + S - 1] ), ordered near -> far by slab
        for (int f = tid; f < MH_F; f += R_THREADS) {
            const uint2 fb = fbin[f];
            if (!(fb.y & 0x80000000u)) continue;
            const int slab = (int)((fb.y >> 20) & 255u);
            const int bx_lo = fb.x & 0xffff, bx_hi = fb.x >> 16, by_lo = fb.y & 1023, by_hi = (fb.y >> 10) & 1023;
            for (int by = by_lo; by <= by_hi; ++by)
                for (int bx = bx_lo; bx <= bx_hi; ++bx) {
                    const int pos = atoms_add(sb + SO_TCOUNT + 4 * ((by * nbx + bx) * S + slab), 1);
                    if (pos < P.bincap) binlist[pos] = (uint16_t)f;
                }
        }
        __syncthreads();
        PROF(1);
        const bool overflow = sint[7] != 0;
        if (overflow && tid == 0) atomicAdd(P.devflags + 1, 1);
        // ---- per-body constants ----
        int q = 0;                                                        // position of person n in the depth order
        for (int k = 0; k < P.N; ++k) if (P.order[t * P.N + k] == n) q = k;
        const uint32_t pre = P.premask[t * P.N + q];
        int sumM = P.rankcnt[t * (P.N + 1) + P.N];
        for (int k = q; k < P.N; ++k) sumM += P.rankcnt[t * (P.N + 1) + k];
        const float Nn = (float)sumM + 1.0f;                              // sum(mask) + 1   (losses.py:36)
        const bool gate = P.mask_valid[t * P.N + q] && P.pose2d_valid[t * P.N + q];      // indexed by POSITION (optimizer.py:472)
        const bool pvalid = P.pose2d_valid[t * P.N + n] != 0;
        const float minz = logf(1.0f + expf(P.zmin_lin[t]));
        const float maxz = minz + 1.0f + logf(1.0f + expf(P.zmax_lin[t]));
        const float izmin = 1.0f / minz, izmax = 1.0f / maxz;
        const float tda = izmin - izmax;
        const size_t plane = (size_t)t * P.H * P.W;
        float aS = 0.f, aA = 0.f, aC = 0.f, aGmin = 0.f, aGmax = 0.f, aSil = 0.f, aCnt = 0.f;
        // ---- P2 / P3: tiles ----
        const int ntiles = overflow ? 0 : ntx * nty;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int ttx = tile % ntx, tty = tile / ntx;
            const int bin = (tty >> ks) * nbx + (ttx >> ks);
            const int off = bin ? tcount[bin * S - 1] : 0, cnt = tcount[bin * S + S - 1] - off;
            if (cnt == 0) continue;
            const int ox = tx0 + ttx * TW, oy = ty0 + tty * TH;           // tile origin (pixels)
            __syncthreads();
            PROF(4);
            // this thread's pixel (the same in the tile set-up and in P3): what the loss needs there.  A pixel that needs no
            // depth fragment (outside the eroded instance mask, optimizer.py:434-438) or no silhouette fragments (hidden by the
            // masks of nearer persons or gated off, :459-475) gets key 0 = "nothing can improve it": P2 prunes every face there
            unsigned pflags;                                              // bit 0 need depth, 1 need silhouette, 2 seg bit, 3 in image
            const int xip = ox + (tid & (TW - 1)), yip = oy + (tid >> 5);
            const bool inimg = (xip < P.W) && (yip < P.H);
            uint32_t cb = 0u, eb = 0u;
            if (MODE == 0 && inimg) {                                     // issued first: the plane loads fly while the descriptors are built
                const size_t pidx = plane + (size_t)yip * P.W + xip;
                cb = P.cbits[pidx]; eb = P.ebits[pidx];
            }
            const int txmax = min(TW, P.W - ox) - 1, tymax = min(TH, P.H - oy) - 1;     // last valid local column / row
            // ---- descriptors of the first chunk of the tile's faces: one thread per face ----
            if (tid < cnt) make_desc(P, sv, binlist[off + tid], ox, oy, txmax, tymax, sdesc + tid * 5);
            pflags = inimg ? 0xbu : 0u;
            if (MODE == 0 && inimg) pflags = 8u | ((pvalid && ((eb >> n) & 1u)) ? 1u : 0u) | ((gate && ((cb & pre) == 0u)) ? 2u : 0u) | (((cb >> n) & 1u) << 2);
            const bool need_d = pflags & 1u, need_s = pflags & 2u;
            dkey[tid] = need_d ? KEY_EMPTY : 0ull;
            if (tid == 0) sint[40] = 0;
#pragma unroll
            for (int s = 0; s < 3; ++s) skey[s * R_THREADS + tid] = KEY_EMPTY;
            skey[3 * R_THREADS + tid] = need_s ? KEY_EMPTY : 0ull;
            if (tid < TW) spx[tid] = (ox + tid < P.W) ? P.pix_x[ox + tid] : 0.f;
            if (tid >= 64 && tid < 64 + TH) spy[tid - 64] = (oy + tid - 64 < P.H) ? P.pix_y[oy + tid - 64] : 0.f;
            const int tile_needed = __syncthreads_or(need_d || need_s);
            PROF(2);
            if (!tile_needed) continue;                                   // uniform: nothing the loss reads in this tile
            // ---- P2: scatter -- each warp takes one face of the chunk at a time (dynamic hand-out) and spreads the face's pixel
            //      rectangle over its lanes, 32 pixels per pass ----
            for (int base = 0; base < cnt; base += R_DESC) {
              if (base > 0) {
                __syncthreads();                                          // everybody is done with the previous chunk
                if (tid == 0) sint[40] = 0;
                if (base + tid < cnt) make_desc(P, sv, binlist[off + base + tid], ox, oy, txmax, tymax, sdesc + tid * 5);
                __syncthreads();
              }
              const int ccnt = min(cnt - base, R_DESC);
              // Two stages per warp.  PRUNE: the current face's pixel rectangle is walked 32 pixels per pass; a pixel survives when the
              // face's nearest vertex is not behind the keys it could displace; survivors (descriptor, pixel) are appended to a
              // per-warp queue (ballot + popc, no divergence).  EVALUATE: as soon as 32 survivors are queued -- from whichever faces --
              // they are evaluated with all lanes busy, each lane loading its own face from the descriptor.
              const uint32_t qa = sb + SO_QUEUE + warp * (R_QUEUE * 4);
              const unsigned ltmask = (1u << lane) - 1u;
              // EVALUATE one queued survivor: entry = pixel (10 bits) | descriptor << 10 | may displace the depth key << 20 | a silhouette key << 21
              auto evaluate = [&](const unsigned ent) {
                    RS_WARP(4);
                    const uint32_t da = sb + SO_SDESC + ((ent >> 10) & 1023u) * 80;
                    const uint32_t ka = sb + 8 * (ent & 1023u);           // + SO_DKEY: depth key of the pixel, + SO_SKEY + s * SK_STRIDE: silhouette keys
                    const bool pd = (ent >> 20) & 1u, ps = (ent >> 21) & 1u;
                    const float4 q0 = lds128<0>(da), q1 = lds128<16>(da), q2 = lds128<32>(da);
                    const float x0 = q0.x, y0 = q0.y, x1 = q0.z, y1 = q0.w, x2 = q1.x, y2 = q1.y, z0 = q1.z, z1 = q1.w, z2 = q2.x;
                    const float inv_den = q2.y;
                    // edge vectors exactly as the oracle rounds them
                    const float ex12 = MH_SUB(x2, x1), ey12 = MH_SUB(y2, y1);
                    const float ex20 = MH_SUB(x0, x2), ey20 = MH_SUB(y0, y2);
                    const float ex01 = MH_SUB(x1, x0), ey01 = MH_SUB(y1, y0);
                    const float px = ldsf<SO_SPX>(sb + 4 * (ent & 31u)), py = ldsf<SO_SPY>(sb + 4 * ((ent >> 5) & 31u));
                    const float dx0 = MH_SUB(px, x0), dy0 = MH_SUB(py, y0);
                    const float dx1 = MH_SUB(px, x1), dy1 = MH_SUB(py, y1);
                    const float dx2 = MH_SUB(px, x2), dy2 = MH_SUB(py, y2);
                    // edge functions exactly as the oracle rounds them (their signs decide `inside`)
                    const float e0 = MH_SUB(MH_MUL(dx1, ey12), MH_MUL(dy1, ex12));
                    const float e1 = MH_SUB(MH_MUL(dx2, ey20), MH_MUL(dy2, ex20));
                    const float e2 = MH_SUB(MH_MUL(dx0, ey01), MH_MUL(dy0, ex01));
                    // depth first: a fragment that cannot displace a key needs no distance test
                    const float c0w = __saturatef(e0 * inv_den), c1w = __saturatef(e1 * inv_den), c2w = __saturatef(e2 * inv_den);
                    const float pz = __fdividef(c0w * z0 + c1w * z1 + c2w * z2, fmaxf(c0w + c1w + c2w, 1e-5f));
                    const float4 q4 = lds128<64>(da);                      // 1/|e12|^2, face
                    const int fcur = __float_as_int(q4.y);
                    // key: depth | face | descriptor -- ordered by depth, then face (the oracle's tie-break); the descriptor index rides along for P3
                    const unsigned long long key = ((unsigned long long)__float_as_uint(pz) << 32) | ((unsigned)fcur << 10) | ((ent >> 10) & 1023u);
                    const bool wd = pd && (pz >= 0.f) && (key < lds64<SO_DKEY>(ka));
                    const bool ws = ps && (pz >= 0.f) && (key < lds64<SO_SKEY + 3 * SK_STRIDE>(ka));
                    if (wd || ws) {
                        RS_ADD(5, 1); RS_WARP(6);
                        const bool dpos = inv_den > 0.f;
                        const bool inside = (e0 != 0.f) && (e1 != 0.f) && (e2 != 0.f) && ((e0 > 0.f) == dpos) && ((e1 > 0.f) == dpos) && ((e2 > 0.f) == dpos);
                        bool vd = inside, vs = inside;
                        if (!inside) {
                            const float d01 = seg_dist_fast(dx0, dy0, ex01, ey01, q2.z, dx1, dy1);
                            const float d02 = seg_dist_fast(dx0, dy0, -ex20, -ey20, q2.w, dx2, dy2);
                            const float d12 = seg_dist_fast(dx1, dy1, ex12, ey12, q4.x, dx2, dy2);
                            const float d = fminf(fminf(d01, d02), d12);
                            vd = d < blur_d_lo; vs = d < blur_s_lo;
                            if (d < blur_d_hi && ((!vd) || (!vs && d < blur_s_hi))) {   // within 1e-5 of a threshold: decide on the exact distance
                                MhFace fc; MhFrag fr;
                                int iv[3];
                                load_face(sv, P.faces, fcur, P.r_d, &fc, iv);
                                mh_face_eval(fc, px, py, &fr);
                                vd = fr.dist < P.blur_d; vs = fr.dist < P.blur_s;
                            }
                            RS_ADD(9, 1);
                        }
                        if (vd && wd) { RS_ADD(7, 1); atoms_min64<SO_DKEY>(ka, key); }
                        if (vs && ws) { RS_ADD(8, 1); key_insert4(ka + SO_SKEY, key); }
                    }
              };
              int qn = 0;                                                 // queued survivors (warp-uniform)
              int kn = 0;
              if (lane == 0) kn = atoms_add(sb + SO_SINT + 4 * 40, MH_R_GRAB);
              for (;;) {
                // next MH_R_GRAB faces of the chunk (dynamic hand-out, one grab ahead; the lists are ordered near -> far, so the coarser
                // grain falls on the cheap, mostly pruned faces at the end)
                const int k2 = __shfl_sync(0xffffffffu, kn, 0);
                if (k2 >= ccnt) break;
                if (lane == 0) kn = atoms_add(sb + SO_SINT + 4 * 40, MH_R_GRAB);
                const int kend2 = min(k2 + MH_R_GRAB, ccnt);
#pragma unroll 1
                for (int k = k2; k < kend2; ++k) {
                const float4 q3 = lds128<48>(sb + SO_SDESC + k * 80);
                const unsigned span = __float_as_uint(q3.x), inner = __float_as_uint(q3.y), zbits = __float_as_uint(q3.z), lanes = __float_as_uint(q3.w);
                if (span == 0u) continue;                                 // binned conservatively, nothing of the face in this tile
                // PRUNE passes over whole rows of the face's rectangle: lane -> (row lr, column) of the first rows once per face,
                // then only the key address advances.  A pixel survives when the face's nearest vertex is not behind the key it could
                // displace (the depth words of the keys alone, conservative on ties); outside the inner rectangle the face cannot be
                // a silhouette fragment at all.  Per-lane bounds fold the lane tests in: a lane beyond the last whole row (or outside
                // the inner columns) has an empty range; every lane loads (addresses stay inside the key planes)
                const unsigned lr = ((unsigned)lane * (lanes & 0x1ffffu)) >> 16;              // lane / w
                const unsigned lrt = lr * ((lanes >> 17) & 63u);                               // lr * (TW - w)
                const bool lane_ok = (unsigned)lane < (lanes >> 23);
                const unsigned pend = (span >> 10) & 0x7ffu, step = span >> 21;
                unsigned p = span & 0x3ffu;                                                    // first pixel of the pass (uniform)
                uint32_t ka = lane_ok ? sb + 8u * (p + (unsigned)lane + lrt) : sb;             // this lane's key slot
                uint32_t kend = lane_ok ? sb + 8u * pend : 0u;
                asm volatile("" : "+r"(kend));                            // keep the per-lane bound in a register (ptxas otherwise re-derives the lane test in every pass)
                const bool col_in = ((unsigned)lane + lrt - (lr << 5)) - ((inner >> 21) & 31u) < (inner >> 26);
                const uint32_t ika0 = sb + 8u * (inner & 0x3ffu);
                uint32_t ikn = (lane_ok && col_in) ? 8u * ((inner >> 10) & 0x7ffu) : 0u;
                asm volatile("" : "+r"(ikn));
                const unsigned ebase = (unsigned)k << 10;
#ifdef MH_RSTATS
                if (lane == 0) { RS_ADD(0, 1); RS_ADD(1, (32u - ((lanes >> 17) & 63u)) * ((pend - (p & ~31u)) >> 5)); }
                int surv_item = 0;
#endif
                for (;;) {
                    // tight loop: passes until the rectangle ends or a batch is full
                    do {
                        RS_WARP(2);
                        const unsigned td = lds32<SO_DKEY + 4>(ka), ts = lds32<SO_SKEY + 3 * SK_STRIDE + 4>(ka);
                        const bool pd = (ka < kend) && (zbits <= td);
                        const bool ps = (ka - ika0 < ikn) && (zbits <= ts);
                        const unsigned bal = __ballot_sync(0xffffffffu, pd || ps);
                        if (pd || ps) {
                            sts32<0>(qa + 4 * (qn + __popc(bal & ltmask)), ebase | ((ka - sb) >> 3) | (pd ? 1u << 20 : 0u) | (ps ? 1u << 21 : 0u));
                            RS_ADD(3, 1);
                        }
                        qn += __popc(bal);
#ifdef MH_RSTATS
                        surv_item += __popc(bal);
#endif
                        ka += 8u * step;
                        p += step;
                    } while (p < pend && qn < 32);
                    if (qn >= 32) {                                       // a full batch
                        __syncwarp();
                        qn -= 32;
                        evaluate(lds32<0>(qa + 4 * (qn + lane)));
                        __syncwarp();
                    }
                    if (p >= pend) break;
                }
#ifdef MH_RSTATS
                if (surv_item == 0 && lane == 0) RS_ADD(10, 1);
#endif
                }
              }
              __syncwarp();
              if (lane < qn) evaluate(lds32<0>(qa + 4 * lane));           // the rest
              __syncwarp();
            }
            __syncthreads();
            PROF(3);
            // ---- P3: one thread per pixel: exact fragments of the winners, losses, silhouette backward ----
            if (!(pflags & 8u)) continue;
            const int xi = ox + (tid & (TW - 1)), yi = oy + (tid >> 5);
            const float pxn = spx[tid & (TW - 1)], pyn = spy[tid >> 5];
            float dz = -1.0f; int df = -1;
            const bool single = cnt <= R_DESC;                           // the tile's descriptors are all still in shared memory
            if ((pflags & 1u) && dkey[tid] != KEY_EMPTY) {
                const unsigned lo = (unsigned)(dkey[tid] & 0xffffffffull);
                df = (int)(lo >> 10);
                float sdu;
                if (single) frag_values_desc(sb + SO_SDESC + (lo & 1023u) * 80, pxn, pyn, &dz, &sdu);
                else frag_values(sv, P.faces, df, pxn, pyn, &dz, &sdu);
            }
            int sf[4]; float sd[4]; float pk[4];
            float prod = 1.0f;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const unsigned long long k = (pflags & 2u) ? skey[s * R_THREADS + tid] : KEY_EMPTY;
                sf[s] = -1; sd[s] = 0.f; pk[s] = 0.f;
                if (k != KEY_EMPTY) {
                    sf[s] = (int)(k & 0xffffffffull);                     // face << 10 | descriptor
                    float pzu;
                    if (single) frag_values_desc(sb + SO_SDESC + (sf[s] & 1023) * 80, pxn, pyn, &pzu, &sd[s]);
                    else frag_values(sv, P.faces, sf[s] >> 10, pxn, pyn, &pzu, &sd[s]);
                    pk[s] = 1.0f / (1.0f + expf(sd[s] / P.sigma));        // sigmoid(-signed / sigma)
                }
                prod = prod * (1.0f - pk[s]);
            }
            const float alpha = 1.0f - prod;
            if (MODE == 1) {
                P.dbg_zbuf[(size_t)yi * P.W + xi] = (df >= 0) ? dz : -1.0f;
                P.dbg_alpha[(size_t)yi * P.W + xi] = alpha;
                continue;
            }
            // ---- depth term (optimizer.py:431-442) ----
            if (df >= 0 && dz > 0.f) {                                  // need_d: pose2d-valid person, pixel inside the eroded mask
                const float zc = fmaxf(dz + 0.2f, P.eps);
                const float zdisp = 1.0f / zc;
                aS += 1.0f;
                aA += logf(fmaxf(zdisp, P.eps));
                const float dd = P.depth[plane + (size_t)yi * P.W + xi];
                const float td = dd * tda + izmax;                        // target_disp (:425)
                aC += logf(fmaxf(td, P.eps));
                if (td >= P.eps) { aGmin += dd / td; aGmax += (1.0f - dd) / td; }
                // d log(clamp(1 / clamp(z + .2, eps), eps)) / dz
                const float gfac = (dz + 0.2f >= P.eps && zdisp >= P.eps) ? -1.0f / zc : 0.f;
                if (gfac != 0.f) {
                    const int wq = atoms_add(sb + SO_SINT + 4 * 6, 1);
                    if (wq < P.wcap) { wpix[wq] = yi * P.W + xi; wface[wq] = df; wz[wq] = gfac; }
                }
            }
            // ---- silhouette term (optimizer.py:459-475, losses.py:35-38) ----
            if (pflags & 2u) {                                           // gate && (cb & pre) == 0
                const float seg = (float)((pflags >> 2) & 1u);
                const float df_ = alpha - seg;
                aSil += df_ * df_;
                aCnt += seg;
                const float ga = P.coef_sil * 2.0f * df_ / Nn;
                if (ga != 0.f) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        if (sf[s] < 0) continue;
                        float others = 1.0f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (j != s) others *= (1.0f - pk[j]);
                        // d alpha / d signed = others * d p / d signed ; p = sigmoid(-signed / sigma)
                        const float gsd = ga * others * (-pk[s] * (1.0f - pk[s]) / P.sigma);
                        const float gdist = sd[s] < 0.f ? -gsd : gsd;     // signed = inside ? -dist : dist
                        if (gdist == 0.f) continue;
                        if (single) sil_grad_desc(sg, sb + SO_SDESC + (sf[s] & 1023) * 80, pxn, pyn, gdist);
                        else sil_grad(sg, sv, P.faces, sf[s] >> 10, pxn, pyn, gdist);
                    }
                }
            }
        }
        __syncthreads();
        PROF(4);
        if (MODE == 1) continue;
        // ---- P4: whole-image sums, then the depth backward over the winner list ----
        aS = wsum(aS); aA = wsum(aA); aC = wsum(aC); aGmin = wsum(aGmin); aGmax = wsum(aGmax); aSil = wsum(aSil); aCnt = wsum(aCnt);
        __syncthreads();
        if (lane == 0) {
            sred[warp] = aS; sred[NW + warp] = aA; sred[2 * NW + warp] = aC; sred[3 * NW + warp] = aGmin; sred[4 * NW + warp] = aGmax;
            sred[5 * NW + warp] = aSil; sred[6 * NW + warp] = aCnt;
        }
        __syncthreads();
        if (tid == 0) {
            float r[7];
            for (int k = 0; k < 7; ++k) { r[k] = 0.f; for (int w = 0; w < NW; ++w) r[k] += sred[NW * k + w]; }
            float* o = P.pfout + (size_t)i * PF_COUNT;
            o[PF_S] = r[0]; o[PF_A] = r[1]; o[PF_C] = r[2]; o[PF_GIZMIN] = r[3]; o[PF_GIZMAX] = r[4];
            // loss = (sum over the whole image of (M (alpha - seg))^2) / Nn ; outside the visited tiles alpha = 0
            const float base = gate ? ((float)P.rankcnt[t * (P.N + 1) + q] - r[6]) : 0.f;
            o[PF_SIL] = gate ? (base + r[5]) / Nn : 0.f;
            o[PF_CNTIN] = r[6];
            const float inv = 1.0f / (r[0] + 1.0f);
            const float diff = r[1] * inv - r[2] * inv;
            o[PF_DEPTHLOSS] = diff * diff;
            sred[7 * NW] = P.coef_depth * 2.0f * diff * inv;              // dL/dA_sum
            if (sint[6] > P.wcap) atomicAdd(P.devflags + 1, 1);
        }
        __syncthreads();
        const float kappa = sred[7 * NW];
        const int nw = min(sint[6], P.wcap);
        if (kappa != 0.f) {
            for (int e = tid; e < nw; e += R_THREADS) {
                const int pix = wpix[e], f = wface[e];
                const int yi = pix / P.W, xi = pix - yi * P.W;
                MhFace fc; int iv[3];
                load_face(sv, P.faces, f, P.r_d, &fc, iv);
                float g[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                mh_face_bwd(fc, P.pix_x[xi], P.pix_y[yi], kappa * wz[e], 0.f, g);
#pragma unroll
                for (int k = 0; k < 9; ++k) if (g[k] != 0.f) grad_add(&sg[3 * iv[k / 3] + (k % 3)], g[k]);
            }
        }
        __syncthreads();
        PROF(5);
        // ---- P5: NDC -> camera-space chain rule, accumulate into dL/dV (this CTA owns the row) ----
        __threadfence();
        __syncthreads();
        const float* vw = P.verts + b * MH_LD3V;
        float* dv = P.dverts + b * MH_LD3V;
        for (int v = tid; v < MH_V; v += R_THREADS) {
            // read at L2 (where the reductions landed), clear for the next body
            const long long qx = __ldcg(&sg[3 * v]), qy = __ldcg(&sg[3 * v + 1]), qz = __ldcg(&sg[3 * v + 2]);
            if (qx != 0) sg[3 * v] = 0;
            if (qy != 0) sg[3 * v + 1] = 0;
            if (qz != 0) sg[3 * v + 2] = 0;
            if ((qx | qy | qz) == 0) continue;
            const float gx = __ll2float_rn(qx) * (1.0f / MH_GFIX), gy = __ll2float_rn(qy) * (1.0f / MH_GFIX), gz = __ll2float_rn(qz) * (1.0f / MH_GFIX);
            const float X = vw[3 * v], Y = vw[3 * v + 1], Z = vw[3 * v + 2];
            const float iz = 1.0f / Z;
            // x_ndc = -k00 X / Z + k02 ; y_ndc = -k11 Y / Z + k12 ; z_view = Z
            dv[3 * v] += -P.k00 * iz * gx;
            dv[3 * v + 1] += -P.k11 * iz * gy;
            dv[3 * v + 2] += (P.k00 * X * gx + P.k11 * Y * gy) * iz * iz + gz;
        }
        if (MODE == 0 && tid == 0) {
            const long long t0c = ((long long)sint[43] << 32) | (unsigned)sint[42];
            P.cost[i] = (int)min(clock64() - t0c, 0x7fffffffll);
        }
        __syncthreads();
        PROF(6);
    }
    RS_FLUSH;
}

// -------------------------------------------------------------------------------------------------
int mh_render_alloc(mh_ctx* c) {
    MhRenderScratch* rs = new MhRenderScratch();
    memset(rs, 0, sizeof(*rs));
    c->rs = rs;
    rs->nctas = c->num_sms;
    rs->maxbins = R_MAXBINS;
    rs->bincap_use = 0; rs->wcap_use = 0;
    rs->bincap = 1 << 20;
    rs->wcap = c->d.H * c->d.W;
    const size_t n = (size_t)rs->nctas;
    cudaError_t e = mh_dev_alloc((void**)&rs->binlist, n * rs->bincap * sizeof(uint16_t));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->fbin, n * MH_F * sizeof(uint2));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->wpix, n * rs->wcap * sizeof(int));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->wface, n * rs->wcap * sizeof(int));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->wz, n * rs->wcap * sizeof(float));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->counter, sizeof(int));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->gsg, n * MH_LD3V * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemset(rs->gsg, 0, n * MH_LD3V * sizeof(long long));
    const size_t tn = (size_t)c->d.T * c->d.N;
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->cost, tn * sizeof(int));
    if (e == cudaSuccess) e = mh_dev_alloc((void**)&rs->work, tn * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(rs->cost, 0, tn * sizeof(int));
    rs->have_cost = false;
    { const char* v = getenv("MH_RENDER_NSLAB"); rs->nslab = v ? std::min(std::max(atoi(v), 1), 256) : R_NSLAB; }     // development switch
    rs->prof = nullptr;
    if (e != cudaSuccess) MH_FAIL(c, MH_E_CUDA, "render scratch: %s", cudaGetErrorString(e));
    rs->smem = (size_t)SO_END + 128;
    e = cudaFuncSetAttribute(k_render<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs->smem);
    if (e != cudaSuccess) MH_FAIL(c, MH_E_CUDA, "render: %zu bytes of shared memory: %s", rs->smem, cudaGetErrorString(e));
    return MH_OK;
}

void mh_render_free(mh_ctx* c) {
    if (!c->rs) return;
    if (c->rs->prof) mh_dev_free(c->rs->prof);
    mh_dev_free(c->rs->fbin); mh_dev_free(c->rs->gsg); mh_dev_free(c->rs->cost); mh_dev_free(c->rs->work);
    mh_dev_free(c->rs->binlist); mh_dev_free(c->rs->wpix); mh_dev_free(c->rs->wface); mh_dev_free(c->rs->wz); mh_dev_free(c->rs->counter);
    delete c->rs;
    c->rs = nullptr;
}

static RenderParams render_params(mh_ctx* c, float blur_d, float blur_s) {
    RenderParams P;
    memset(&P, 0, sizeof(P));
    const mh_dims& d = c->d;
    P.verts = c->verts; P.dverts = c->dverts; P.faces = c->faces; P.pix_x = c->pix_x; P.pix_y = c->pix_y;
    P.depth = c->depth; P.cbits = c->cbits; P.ebits = c->ebits; P.order = c->order; P.premask = c->premask; P.rankcnt = c->rankcnt;
    P.pose2d_valid = c->pose2d_valid; P.mask_valid = c->mask_valid;
    P.zmin_lin = c->params + c->off[MH_P_ZMIN_LIN]; P.zmax_lin = c->params + c->off[MH_P_ZMAX_LIN];
    P.pfout = c->pfout; P.devflags = c->devflags;
    P.binlist = c->rs->binlist; P.bincap = c->rs->bincap; P.fbin = c->rs->fbin; P.wpix = c->rs->wpix; P.wface = c->rs->wface; P.wz = c->rs->wz; P.wcap = c->rs->wcap;
    P.counter = c->rs->counter; P.gsg = c->rs->gsg; P.nslab = c->rs->nslab;
    P.cost = c->rs->cost; P.work = nullptr;
    P.maxbins = c->rs->maxbins;
    if (c->rs->bincap_use) P.bincap = c->rs->bincap_use;
    if (c->rs->wcap_use) P.wcap = c->rs->wcap_use;
    P.prof = c->rs->prof;
    P.T = d.T; P.N = d.N; P.H = d.H; P.W = d.W;
    P.k00 = c->Kndc[0]; P.k02 = c->Kndc[2]; P.k11 = c->Kndc[5]; P.k12 = c->Kndc[6];
    P.rx = d.W > d.H ? (float)(2.0 * d.W / d.H) : 2.0f;
    P.ry = d.H > d.W ? (float)(2.0 * d.H / d.W) : 2.0f;
    P.blur_d = blur_d; P.blur_s = blur_s; P.r_d = sqrtf(blur_d > blur_s ? blur_d : blur_s);
    P.r_s = fminf(sqrtf(blur_s) * 1.001f, P.r_d);
    P.sigma = 1e-4f;                 // BlendParams default used by SoftSilhouetteShader
    P.eps = c->c.eps;
    P.coef_depth = c->c.depth; P.coef_sil = c->c.silhouette;
    return P;
}

// hand-out order of the next launch: bodies by the cycles they took in the last one, longest first (ties by index).  Rank sort: every
// thread counts the bodies ahead of its own -- n^2 compares, 17 M at C3, spread over n threads
__global__ void __launch_bounds__(256) k_render_rank(const int* __restrict__ cost, int n, int* __restrict__ work) {
    __shared__ int sc[256];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int ci = i < n ? cost[i] : 0;
    int rank = 0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        __syncthreads();
        sc[threadIdx.x] = (j0 + threadIdx.x < n) ? cost[j0 + threadIdx.x] : -1;
        __syncthreads();
        const int m = min(256, n - j0);
        for (int k = 0; k < m; ++k) { const int cj = sc[k]; rank += (cj > ci) || (cj == ci && j0 + k < i); }
    }
    if (i < n) work[rank] = i;
}

int mh_render_all(mh_ctx* c, cudaStream_t st) {
    RenderParams P = render_params(c, 1e-4f, 2e-5f);          // optimizer.py:213, 223
    const int TNb = c->d.T * c->d.N;
    static const int use_cost = [] { const char* v = getenv("MH_RENDER_COST_ORDER"); return v ? atoi(v) : 1; }();
    if (use_cost && c->rs->have_cost) {
        // bodies differ 10x in projected size and a rank holds only a few per CTA when the frames are sharded: the measured cost of
        // the previous cycle is a far better predictor of this cycle's than the depth rank alone
        k_render_rank<<<mh_cdiv(TNb, 256), 256, 0, st>>>(c->rs->cost, TNb, c->rs->work);
        MH_LAUNCHED(c);
        P.work = c->rs->work;
    }
    c->rs->have_cost = true;
    MH_CUDA(c, cudaMemsetAsync(c->rs->counter, 0, sizeof(int), st));
    const int grid = std::min(c->rs->nctas, c->d.T * c->d.N);
    k_render<0><<<grid, R_THREADS, c->rs->smem, st>>>(P);
    MH_LAUNCHED(c);
    return MH_OK;
}

__global__ void k_fill2(float* a, float va, float* b, float vb, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { a[i] = va; b[i] = vb; }
}

int mh_render_planes(mh_ctx* c, int t, int n, float* zbuf_dev, float* alpha_dev, float blur_d, float blur_s, cudaStream_t st) {
    RenderParams P = render_params(c, blur_d, blur_s);
    P.dbg_zbuf = zbuf_dev; P.dbg_alpha = alpha_dev; P.dbg_body = t * c->d.N + n;
    const int64_t HW = (int64_t)c->d.H * c->d.W;
    k_fill2<<<mh_cdiv(HW, 1024), 256, 0, st>>>(zbuf_dev, -1.0f, alpha_dev, 0.0f, HW);
    MH_LAUNCHED(c);
    k_render<1><<<1, R_THREADS, c->rs->smem, st>>>(P);
    MH_LAUNCHED(c);
    return MH_OK;
}

int mh_render_debug(mh_ctx* c, int t, int n, float* zbuf_dev, float* alpha_dev, cudaStream_t st) {
    MH_TRY(mh_forward_only(c, st));
    return mh_render_planes(c, t, n, zbuf_dev, alpha_dev, 1e-4f, 2e-5f, st);
}

// Development aid: per-phase cycle counters of the render kernel, summed over the CTAs (8 slots:
// load+ndc | binning | tile set-up + descriptors | prune + evaluate | per-pixel + silhouette backward | sums + depth backward | chain rule | -),
// then the pair statistics of an -DMH_RSTATS build.
extern "C" int mh_render_profile(mh_ctx* c, int32_t on, long long* out32_host) {
    if (!c || !c->rs) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    MhRenderScratch* rs = c->rs;
    const size_t n = (size_t)rs->nctas * MH_NPROF;
    if (out32_host && rs->prof) {
        std::vector<long long> h(n);
        MH_CUDA(c, cudaDeviceSynchronize());
        MH_CUDA(c, cudaMemcpy(h.data(), rs->prof, n * sizeof(long long), cudaMemcpyDeviceToHost));
        for (int k = 0; k < MH_NPROF; ++k) { out32_host[k] = 0; for (int b = 0; b < rs->nctas; ++b) out32_host[k] += h[(size_t)b * MH_NPROF + k]; }
    }
    if (on && !rs->prof) MH_CUDA(c, mh_dev_alloc((void**)&rs->prof, n * sizeof(long long)));
    if (on) MH_CUDA(c, cudaMemset(rs->prof, 0, n * sizeof(long long)));
    if (!on && rs->prof) { mh_dev_free(rs->prof); rs->prof = nullptr; }
    return MH_OK;
}

// Testing aid: shrink the render capacities so that small inputs exercise the coarse-binning path (maxbins) and the
// capacity-error paths (tile-list entries, depth-winner entries).  0 keeps a value.
extern "C" int mh_debug_set_render_caps(mh_ctx* c, int32_t maxbins, int32_t bincap, int32_t wcap) {
    if (!c || !c->rs) return MH_E_ARG;
    MhRenderScratch* rs = c->rs;
    if (maxbins < 0 || maxbins > R_MAXBINS || bincap < 0 || bincap > (1 << 20) || wcap < 0 || wcap > c->d.H * c->d.W)
        MH_FAIL(c, MH_E_ARG, "mh_debug_set_render_caps: capacities can only be reduced");
    if (maxbins) rs->maxbins = maxbins;
    if (bincap) rs->bincap_use = bincap;
    if (wcap) rs->wcap_use = wcap;
    cudaSetDevice(c->d.device);
    MH_CUDA(c, cudaMemset(c->devflags + 1, 0, sizeof(int)));
    return MH_OK;
}
