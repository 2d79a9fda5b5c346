// Small fixed-size math shared by the kernels (and compiled for the host by
// tests/hostmath to check the analytic derivatives on CPU against autograd).
//
// Reference semantics followed (paths relative to the reference repo):
//   rodrigues        mhmocap/smpl.py:647-678   (angle = ||r + 1e-8||, axis = r / angle)
//   kinematic chain  mhmocap/smpl.py:692-746
//   projection       mhmocap/transforms.py:57-95 (incl. the code's own distortion form)
//   face evaluation  PyTorch3D naive rasteriser semantics, SURVEY.md Appendix A
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MH_HD __host__ __device__ __forceinline__
#else
#define MH_HD inline
#endif

#define MH_NJ 24          // SMPL joints
#define MH_NJR 17         // regressed (AlphaPose) joints
#define MH_NPOSE 72
#define MH_NBETA 10
#define MH_NPF 207        // pose-feature length
#define MH_NPF_LIVE 189   // joints 22, 23 are forced to identity -> their 18 features are exactly 0
#define MH_KEPS 1e-8f

// parents of the SMPL kinematic tree (smpl.py:269-272)
#define MH_PARENTS {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21}

// ---------------------------------------------------------------------------
// Rodrigues
// ---------------------------------------------------------------------------
MH_HD void mh_rodrigues(const float r[3], float R[9]) {
    const float ux = r[0] + 1e-8f, uy = r[1] + 1e-8f, uz = r[2] + 1e-8f;
    const float a = sqrtf(ux * ux + uy * uy + uz * uz);
    const float kx = r[0] / a, ky = r[1] / a, kz = r[2] / a;
    float s, c;
#ifdef __CUDA_ARCH__
    sincosf(a, &s, &c);
#else
    s = sinf(a); c = cosf(a);
#endif
    const float c1 = 1.0f - c;
    const float kk = kx * kx + ky * ky + kz * kz;
    // K^2 = k k^T - (k.k) I   (|k| is not exactly 1 because of the 1e-8 offset)
    R[0] = 1.0f + c1 * (kx * kx - kk);  R[1] = -s * kz + c1 * kx * ky;       R[2] = s * ky + c1 * kx * kz;
    R[3] = s * kz + c1 * kx * ky;       R[4] = 1.0f + c1 * (ky * ky - kk);   R[5] = -s * kx + c1 * ky * kz;
    R[6] = -s * ky + c1 * kx * kz;      R[7] = s * kx + c1 * ky * kz;        R[8] = 1.0f + c1 * (kz * kz - kk);
}

// gradient of a scalar w.r.t. r given G = dL/dR (row-major 3x3)
MH_HD void mh_rodrigues_bwd(const float r[3], const float G[9], float gr[3]) {
    const float ux = r[0] + 1e-8f, uy = r[1] + 1e-8f, uz = r[2] + 1e-8f;
    const float a = sqrtf(ux * ux + uy * uy + uz * uz);
    const float ia = 1.0f / a;
    const float kx = r[0] * ia, ky = r[1] * ia, kz = r[2] * ia;
    float s, c;
#ifdef __CUDA_ARCH__
    sincosf(a, &s, &c);
#else
    s = sinf(a); c = cosf(a);
#endif
    const float c1 = 1.0f - c;
    const float kk = kx * kx + ky * ky + kz * kz;
    const float trG = G[0] + G[4] + G[8];
    // <G, K> and <G, K^2>
    const float gK = kx * (G[7] - G[5]) + ky * (G[2] - G[6]) + kz * (G[3] - G[1]);
    const float Gk0 = G[0] * kx + G[1] * ky + G[2] * kz;
    const float Gk1 = G[3] * kx + G[4] * ky + G[5] * kz;
    const float Gk2 = G[6] * kx + G[7] * ky + G[8] * kz;
    const float Gtk0 = G[0] * kx + G[3] * ky + G[6] * kz;
    const float Gtk1 = G[1] * kx + G[4] * ky + G[7] * kz;
    const float Gtk2 = G[2] * kx + G[5] * ky + G[8] * kz;
    const float gK2 = (kx * Gk0 + ky * Gk1 + kz * Gk2) - kk * trG;
    const float ga = c * gK + s * gK2;
    // dL/dk
    const float gk0 = s * (G[7] - G[5]) + c1 * (Gk0 + Gtk0 - 2.0f * trG * kx);
    const float gk1 = s * (G[2] - G[6]) + c1 * (Gk1 + Gtk1 - 2.0f * trG * ky);
    const float gk2 = s * (G[3] - G[1]) + c1 * (Gk2 + Gtk2 - 2.0f * trG * kz);
    // k = r / a, a = |u|:  dk_i/dr_m = delta_im / a - r_i u_m / a^3 ;  da/dr_m = u_m / a
    const float gkr = gk0 * r[0] + gk1 * r[1] + gk2 * r[2];
    const float t = ga * ia - gkr * ia * ia * ia;
    gr[0] = gk0 * ia + t * ux;
    gr[1] = gk1 * ia + t * uy;
    gr[2] = gk2 * ia + t * uz;
}

// ---------------------------------------------------------------------------
// 3x3 helpers (row-major)
// ---------------------------------------------------------------------------
MH_HD void mh_mat3_mul(const float A[9], const float B[9], float C[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
MH_HD void mh_mat3_vec(const float A[9], const float v[3], float o[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
MH_HD void mh_mat3t_vec(const float A[9], const float v[3], float o[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}

// ---------------------------------------------------------------------------
// Pose stage for one body: theta (72), rest joints J (24x3) ->
//   A (24 x 12: rows [R | t]), pose feature pf (207), posed joints optional.
// Follows lbs (smpl.py:541-547) and batch_rigid_transform (smpl.py:716-746).
// ---------------------------------------------------------------------------
MH_HD void mh_pose_forward(const float* theta, const float* J, float* A /*24*12*/, float* pf /*207 or null*/,
                           float* Rout /*24*9 or null*/) {
    const int parents[MH_NJ] = MH_PARENTS;
    float GR[MH_NJ][9];
    float Gt[MH_NJ][3];
    for (int j = 0; j < MH_NJ; ++j) {
        float R[9];
        if (j < 22) {
            mh_rodrigues(theta + 3 * j, R);
        } else {
            R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        }
        if (Rout) for (int e = 0; e < 9; ++e) Rout[j * 9 + e] = R[e];
        if (pf && j >= 1) {
            for (int e = 0; e < 9; ++e) pf[(j - 1) * 9 + e] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
        }
        if (j == 0) {
            for (int e = 0; e < 9; ++e) GR[0][e] = R[e];
            Gt[0][0] = J[0]; Gt[0][1] = J[1]; Gt[0][2] = J[2];
        } else {
            const int p = parents[j];
            mh_mat3_mul(GR[p], R, GR[j]);
            float rel[3] = {J[3 * j] - J[3 * p], J[3 * j + 1] - J[3 * p + 1], J[3 * j + 2] - J[3 * p + 2]};
            float o[3];
            mh_mat3_vec(GR[p], rel, o);
            Gt[j][0] = o[0] + Gt[p][0]; Gt[j][1] = o[1] + Gt[p][1]; Gt[j][2] = o[2] + Gt[p][2];
        }
        float rj[3];
        mh_mat3_vec(GR[j], J + 3 * j, rj);
        float* a = A + j * 12;
        a[0] = GR[j][0]; a[1] = GR[j][1]; a[2] = GR[j][2];  a[3] = Gt[j][0] - rj[0];
        a[4] = GR[j][3]; a[5] = GR[j][4]; a[6] = GR[j][5];  a[7] = Gt[j][1] - rj[1];
        a[8] = GR[j][6]; a[9] = GR[j][7]; a[10] = GR[j][8]; a[11] = Gt[j][2] - rj[2];
    }
}

// Backward of the pose stage for one body.
//   in : theta (72), J (24x3), dA (24x12, gradient w.r.t. A), dpf (207, gradient w.r.t. pose feature)
//   out: dtheta (72) (entries 66..71 = 0), dJ (24x3)
MH_HD void mh_pose_backward(const float* theta, const float* J, const float* dA, const float* dpf,
                            float* dtheta, float* dJ) {
    const int parents[MH_NJ] = MH_PARENTS;
    float R[MH_NJ][9];
    float GR[MH_NJ][9];
    for (int j = 0; j < MH_NJ; ++j) {
        if (j < 22) mh_rodrigues(theta + 3 * j, R[j]);
        else { for (int e = 0; e < 9; ++e) R[j][e] = (e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f; }
        if (j == 0) { for (int e = 0; e < 9; ++e) GR[0][e] = R[0][e]; }
        else mh_mat3_mul(GR[parents[j]], R[j], GR[j]);
    }
    float dGR[MH_NJ][9];
    float dGt[MH_NJ][3];
    for (int j = 0; j < MH_NJ; ++j) {
        const float* a = dA + j * 12;
        const float dt[3] = {a[3], a[7], a[11]};
        // A.R = G.R ; A.t = G.t - G.R J   =>  dG.R = dA.R - dA.t J^T ; dG.t = dA.t ; dJ = -G.R^T dA.t
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k) dGR[j][i * 3 + k] = a[i * 4 + k] - dt[i] * J[3 * j + k];
        dGt[j][0] = dt[0]; dGt[j][1] = dt[1]; dGt[j][2] = dt[2];
        float o[3];
        mh_mat3t_vec(GR[j], dt, o);
        dJ[3 * j] = -o[0]; dJ[3 * j + 1] = -o[1]; dJ[3 * j + 2] = -o[2];
    }
    for (int j = MH_NJ - 1; j >= 0; --j) {
        float dR[9];
        if (j == 0) {
            for (int e = 0; e < 9; ++e) dR[e] = dGR[0][e];
            dJ[0] += dGt[0][0]; dJ[1] += dGt[0][1]; dJ[2] += dGt[0][2];
        } else {
            const int p = parents[j];
            // G_j.R = G_p.R R_j
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) {
                    // dG_p.R += dG_j.R R_j^T
                    dGR[p][i * 3 + k] += dGR[j][i * 3] * R[j][k * 3] + dGR[j][i * 3 + 1] * R[j][k * 3 + 1] + dGR[j][i * 3 + 2] * R[j][k * 3 + 2];
                    // dR_j = G_p.R^T dG_j.R
                    dR[i * 3 + k] = GR[p][i] * dGR[j][k] + GR[p][3 + i] * dGR[j][3 + k] + GR[p][6 + i] * dGR[j][6 + k];
                }
            // G_j.t = G_p.R rel_j + G_p.t
            const float rel[3] = {J[3 * j] - J[3 * p], J[3 * j + 1] - J[3 * p + 1], J[3 * j + 2] - J[3 * p + 2]};
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) dGR[p][i * 3 + k] += dGt[j][i] * rel[k];
            float drel[3];
            mh_mat3t_vec(GR[p], dGt[j], drel);
            for (int k = 0; k < 3; ++k) {
                dGt[p][k] += dGt[j][k];
                dJ[3 * j + k] += drel[k];
                dJ[3 * p + k] -= drel[k];
            }
        }
        if (j >= 1 && dpf) for (int e = 0; e < 9; ++e) dR[e] += dpf[(j - 1) * 9 + e];
        if (j < 22) mh_rodrigues_bwd(theta + 3 * j, dR, dtheta + 3 * j);
        else { dtheta[3 * j] = 0; dtheta[3 * j + 1] = 0; dtheta[3 * j + 2] = 0; }
    }
}

// ---------------------------------------------------------------------------
// Camera projection of one point (transforms.py:57-95).  K = 3x3 row-major,
// Kd = 5 distortion coefficients or null.  uv out; optional gradient:
// given (gu, gv) returns dL/dP.
// ---------------------------------------------------------------------------
MH_HD void mh_project(const float P[3], const float* K, const float* Kd, float uv[2]) {
    float x = P[0] / P[2], y = P[1] / P[2];
    if (Kd) {
        const float r = x * x + y * y;
        const float rad = 1.0f + Kd[0] * r + Kd[1] * r * r + Kd[4] * r * r * r;
        const float xx = x * rad + 2.0f * Kd[2] * x * y + Kd[3] * (r + 2.0f * x * x);
        const float yy = y * rad + 2.0f * Kd[3] * y * y + Kd[2] * (r + 2.0f * y * y);
        x = xx; y = yy;
    }
    uv[0] = K[0] * x + K[1] * y + K[2];
    uv[1] = K[3] * x + K[4] * y + K[5];
}

MH_HD void mh_project_bwd(const float P[3], const float* K, const float* Kd, float gu, float gv, float gP[3]) {
    const float iz = 1.0f / P[2];
    const float x = P[0] * iz, y = P[1] * iz;
    float gxx = K[0] * gu + K[3] * gv;      // dL/d(distorted x)
    float gyy = K[1] * gu + K[4] * gv;
    float gx = gxx, gy = gyy;
    if (Kd) {
        const float r = x * x + y * y;
        const float rad = 1.0f + Kd[0] * r + Kd[1] * r * r + Kd[4] * r * r * r;
        const float drad = Kd[0] + 2.0f * Kd[1] * r + 3.0f * Kd[4] * r * r;     // d rad / d r
        // xx = x rad + 2 p1 x y + p2 (r + 2 x^2) ; yy = y rad + 2 p2 y^2 + p1 (r + 2 y^2)   (p1 = Kd[2], p2 = Kd[3])
        const float dxx_dx = rad + x * drad * 2.0f * x + 2.0f * Kd[2] * y + Kd[3] * (2.0f * x + 4.0f * x);
        const float dxx_dy = x * drad * 2.0f * y + 2.0f * Kd[2] * x + Kd[3] * (2.0f * y);
        const float dyy_dx = y * drad * 2.0f * x + Kd[2] * (2.0f * x);
        const float dyy_dy = rad + y * drad * 2.0f * y + 4.0f * Kd[3] * y + Kd[2] * (2.0f * y + 4.0f * y);
        gx = gxx * dxx_dx + gyy * dyy_dx;
        gy = gxx * dxx_dy + gyy * dyy_dy;
    }
    gP[0] = gx * iz;
    gP[1] = gy * iz;
    gP[2] = -(x * gx + y * gy) * iz;
}

// ---------------------------------------------------------------------------
// Face evaluation (PyTorch3D naive semantics, Appendix A of SURVEY.md).
//
// The FORWARD arithmetic mirrors oracle/raster.py operation by operation (every
// product, sum and quotient individually rounded, no FMA contraction, true
// divisions) so that the discrete decisions -- inside test, blur-radius test,
// nearest-first ordering -- are bit-identical to the oracle's.  MH_MUL / MH_ADD /
// MH_SUB / MH_DIV are the non-contractable forms on the device; the host build
// is compiled with -ffp-contract=off.
// ---------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define MH_MUL(a, b) __fmul_rn((a), (b))
#define MH_ADD(a, b) __fadd_rn((a), (b))
#define MH_SUB(a, b) __fsub_rn((a), (b))
#define MH_DIV(a, b) __fdiv_rn((a), (b))
#else
#define MH_MUL(a, b) ((a) * (b))
#define MH_ADD(a, b) ((a) + (b))
#define MH_SUB(a, b) ((a) - (b))
#define MH_DIV(a, b) ((a) / (b))
#endif

// world (camera-space metres) -> (x_ndc, y_ndc, z_view): R = diag(-1,-1,1), T = 0, then the NDC calibration
// (optimizer.py:204-207; oracle.raster.world_to_ndc)
MH_HD void mh_world_to_ndc(const float P[3], float k00, float k02, float k11, float k12, float o[3]) {
    const float xv = -P[0], yv = -P[1], zv = P[2];
    o[0] = MH_DIV(MH_ADD(MH_MUL(k00, xv), MH_MUL(k02, zv)), zv);
    o[1] = MH_DIV(MH_ADD(MH_MUL(k11, yv), MH_MUL(k12, zv)), zv);
    o[2] = zv;
}

struct MhEdge { float bax, bay, safe; int deg; };     // b - a, |b - a|^2 (1 when degenerate), degenerate flag

struct MhFace {
    float x0, y0, x1, y1, x2, y2, z0, z1, z2;
    float den;                 // area + 1e-8
    MhEdge e01, e02, e12;
    float xmin, xmax, ymin, ymax;   // bbox inflated by sqrt(blur radius) of the DEPTH raster
    int skip;                  // max z < 0 or |area| <= 1e-8
};

MH_HD float mh_edge(float px, float py, float ax, float ay, float bx, float by) {
    return MH_SUB(MH_MUL(MH_SUB(px, ax), MH_SUB(by, ay)), MH_MUL(MH_SUB(py, ay), MH_SUB(bx, ax)));
}

MH_HD void mh_edge_setup(float ax, float ay, float bx, float by, MhEdge* e) {
    e->bax = MH_SUB(bx, ax); e->bay = MH_SUB(by, ay);
    const float l2 = MH_ADD(MH_MUL(e->bax, e->bax), MH_MUL(e->bay, e->bay));
    e->deg = l2 <= MH_KEPS;
    e->safe = e->deg ? 1.0f : l2;
}

MH_HD void mh_face_setup(const float v0[3], const float v1[3], const float v2[3], float r_inflate, MhFace* f) {
    f->x0 = v0[0]; f->y0 = v0[1]; f->x1 = v1[0]; f->y1 = v1[1]; f->x2 = v2[0]; f->y2 = v2[1];
    f->z0 = v0[2]; f->z1 = v1[2]; f->z2 = v2[2];
    const float area = mh_edge(v2[0], v2[1], v0[0], v0[1], v1[0], v1[1]);
    f->den = MH_ADD(area, MH_KEPS);
    mh_edge_setup(v0[0], v0[1], v1[0], v1[1], &f->e01);
    mh_edge_setup(v0[0], v0[1], v2[0], v2[1], &f->e02);
    mh_edge_setup(v1[0], v1[1], v2[0], v2[1], &f->e12);
    f->xmin = MH_SUB(fminf(fminf(v0[0], v1[0]), v2[0]), r_inflate);
    f->xmax = MH_ADD(fmaxf(fmaxf(v0[0], v1[0]), v2[0]), r_inflate);
    f->ymin = MH_SUB(fminf(fminf(v0[1], v1[1]), v2[1]), r_inflate);
    f->ymax = MH_ADD(fmaxf(fmaxf(v0[1], v1[1]), v2[1]), r_inflate);
    const float zmax = fmaxf(v0[2], fmaxf(v1[2], v2[2]));
    const bool zero_area = (area <= MH_KEPS) && (area >= -MH_KEPS);
    f->skip = (zmax < 0.0f || zero_area || !(zmax == zmax)) ? 1 : 0;
}

MH_HD float mh_clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// squared distance from p to segment (a, b) (degenerate segments: distance to endpoint b, as upstream);
// t_out = clamped parameter (1 when degenerate).
MH_HD float mh_seg_dist(float px, float py, float ax, float ay, float bx, float by, const MhEdge& e, float* t_out) {
    if (e.deg) {
        *t_out = 1.0f;
        const float dx = MH_SUB(px, bx), dy = MH_SUB(py, by);
        return MH_ADD(MH_MUL(dx, dx), MH_MUL(dy, dy));
    }
    float t = MH_DIV(MH_ADD(MH_MUL(e.bax, MH_SUB(px, ax)), MH_MUL(e.bay, MH_SUB(py, ay))), e.safe);
    t = mh_clamp01(t);
    const float qx = MH_ADD(ax, MH_MUL(t, e.bax)), qy = MH_ADD(ay, MH_MUL(t, e.bay));
    const float dx = MH_SUB(px, qx), dy = MH_SUB(py, qy);
    *t_out = t;
    return MH_ADD(MH_MUL(dx, dx), MH_MUL(dy, dy));
}

struct MhFrag {
    float pz;        // clipped-barycentric depth
    float dist;      // unsigned squared distance to the nearest edge
    bool inside;
};

// false when the pixel is outside the inflated bbox or the clipped depth is negative
MH_HD bool mh_face_eval(const MhFace& f, float px, float py, MhFrag* o) {
    if (px > f.xmax || px < f.xmin || py > f.ymax || py < f.ymin) return false;
    const float w0 = MH_DIV(mh_edge(px, py, f.x1, f.y1, f.x2, f.y2), f.den);
    const float w1 = MH_DIV(mh_edge(px, py, f.x2, f.y2, f.x0, f.y0), f.den);
    const float w2 = MH_DIV(mh_edge(px, py, f.x0, f.y0, f.x1, f.y1), f.den);
    float c0 = mh_clamp01(w0), c1 = mh_clamp01(w1), c2 = mh_clamp01(w2);
    const float bs = fmaxf(MH_ADD(MH_ADD(c0, c1), c2), 1e-5f);
    c0 = MH_DIV(c0, bs); c1 = MH_DIV(c1, bs); c2 = MH_DIV(c2, bs);
    const float pz = MH_ADD(MH_ADD(MH_MUL(c0, f.z0), MH_MUL(c1, f.z1)), MH_MUL(c2, f.z2));
    float t;
    const float d01 = mh_seg_dist(px, py, f.x0, f.y0, f.x1, f.y1, f.e01, &t);
    const float d02 = mh_seg_dist(px, py, f.x0, f.y0, f.x2, f.y2, f.e02, &t);
    const float d12 = mh_seg_dist(px, py, f.x1, f.y1, f.x2, f.y2, f.e12, &t);
    o->pz = pz;
    o->dist = fminf(fminf(d01, d02), d12);
    o->inside = (w0 > 0.0f) && (w1 > 0.0f) && (w2 > 0.0f);
    return pz >= 0.0f;
}

// Backward of one fragment.  gz = dL/d(pz) (0 if unused), gd = dL/d(UNSIGNED dist) (0 if unused).
// Accumulates into g[9] = d/d(x0,y0,z0,x1,y1,z1,x2,y2,z2) (NDC xy, view z).
MH_HD void mh_face_bwd(const MhFace& f, float px, float py, float gz, float gd, float g[9]) {
    if (gz != 0.0f) {
        const float inv_den = 1.0f / f.den;
        const float e0 = mh_edge(px, py, f.x1, f.y1, f.x2, f.y2);
        const float e1 = mh_edge(px, py, f.x2, f.y2, f.x0, f.y0);
        const float e2 = mh_edge(px, py, f.x0, f.y0, f.x1, f.y1);
        const float w[3] = {e0 * inv_den, e1 * inv_den, e2 * inv_den};
        float c[3];
        for (int i = 0; i < 3; ++i) c[i] = mh_clamp01(w[i]);
        const float sum = c[0] + c[1] + c[2];
        const float bs = fmaxf(sum, 1e-5f);
        const float z[3] = {f.z0, f.z1, f.z2};
        float dc[3];
        float dbs = 0.0f;
        for (int i = 0; i < 3; ++i) {
            g[3 * i + 2] += gz * c[i] / bs;             // d pz / d z_i = clipped normalised barycentric
            dc[i] = gz * z[i] / bs;
            dbs -= gz * z[i] * c[i] / (bs * bs);
        }
        if (sum >= 1e-5f) for (int i = 0; i < 3; ++i) dc[i] += dbs;
        float dw[3];
        for (int i = 0; i < 3; ++i) dw[i] = (w[i] >= 0.0f && w[i] <= 1.0f) ? dc[i] : 0.0f;
        // w_i = e_i / den
        const float de0 = dw[0] * inv_den, de1 = dw[1] * inv_den, de2 = dw[2] * inv_den;
        const float dden = -(dw[0] * w[0] + dw[1] * w[1] + dw[2] * w[2]) * inv_den;
        // e(p,a,b): d/dax = py - by ; d/day = bx - px ; d/dbx = -(py - ay) ; d/dby = px - ax
        // e0 = e(p, v1, v2)
        g[3] += de0 * (py - f.y2); g[4] += de0 * (f.x2 - px); g[6] += de0 * -(py - f.y1); g[7] += de0 * (px - f.x1);
        // e1 = e(p, v2, v0)
        g[6] += de1 * (py - f.y0); g[7] += de1 * (f.x0 - px); g[0] += de1 * -(py - f.y2); g[1] += de1 * (px - f.x2);
        // e2 = e(p, v0, v1)
        g[0] += de2 * (py - f.y1); g[1] += de2 * (f.x1 - px); g[3] += de2 * -(py - f.y0); g[4] += de2 * (px - f.x0);
        // area = e(v2, v0, v1): d/dv2 = (y1 - y0, -(x1 - x0)) ; d/dv0 = (y2 - y1, x1 - x2) ; d/dv1 = (-(y2 - y0), x2 - x0)
        g[6] += dden * (f.y1 - f.y0); g[7] += dden * -(f.x1 - f.x0);
        g[0] += dden * (f.y2 - f.y1); g[1] += dden * (f.x1 - f.x2);
        g[3] += dden * -(f.y2 - f.y0); g[4] += dden * (f.x2 - f.x0);
    }
    if (gd != 0.0f) {
        float t01, t02, t12;
        const float d01 = mh_seg_dist(px, py, f.x0, f.y0, f.x1, f.y1, f.e01, &t01);
        const float d02 = mh_seg_dist(px, py, f.x0, f.y0, f.x2, f.y2, f.e02, &t02);
        const float d12 = mh_seg_dist(px, py, f.x1, f.y1, f.x2, f.y2, f.e12, &t12);
        // nearest edge (first minimum in the order 01, 02, 12)
        int ia, ib; float ax, ay, bx, by, t;
        if (d01 <= d02 && d01 <= d12) { ia = 0; ib = 1; ax = f.x0; ay = f.y0; bx = f.x1; by = f.y1; t = t01; }
        else if (d02 <= d12)          { ia = 0; ib = 2; ax = f.x0; ay = f.y0; bx = f.x2; by = f.y2; t = t02; }
        else                          { ia = 1; ib = 2; ax = f.x1; ay = f.y1; bx = f.x2; by = f.y2; t = t12; }
        const float qx = px - (ax + t * (bx - ax)), qy = py - (ay + t * (by - ay));
        // d = |p - a - t (b - a)|^2 with t at its (clamped) optimum: dd/da = -2 q (1 - t), dd/db = -2 q t
        g[3 * ia] += gd * -2.0f * qx * (1.0f - t);  g[3 * ia + 1] += gd * -2.0f * qy * (1.0f - t);
        g[3 * ib] += gd * -2.0f * qx * t;           g[3 * ib + 1] += gd * -2.0f * qy * t;
    }
}
