// Host build of the kernels' shared math header (TEST ONLY: lets the CPU test-suite check the
// analytic derivatives in mh_math.cuh against torch autograd without a GPU; the product never loads this).
#include "../../scene-aware-3d-multi-human_b200/csrc/mh_math.cuh"
extern "C" {
void hm_rodrigues(const float* r, float* R) { mh_rodrigues(r, R); }
void hm_rodrigues_bwd(const float* r, const float* G, float* gr) { mh_rodrigues_bwd(r, G, gr); }
void hm_pose_forward(const float* theta, const float* J, float* A, float* pf, float* R) { mh_pose_forward(theta, J, A, pf, R); }
void hm_pose_backward(const float* theta, const float* J, const float* dA, const float* dpf, float* dtheta, float* dJ) {
    mh_pose_backward(theta, J, dA, dpf, dtheta, dJ);
}
void hm_project(const float* P, const float* K, const float* Kd, float* uv) { mh_project(P, K, Kd, uv); }
void hm_project_bwd(const float* P, const float* K, const float* Kd, float gu, float gv, float* gP) { mh_project_bwd(P, K, Kd, gu, gv, gP); }
// verts: 9 floats (x0,y0,z0,x1,y1,z1,x2,y2,z2) NDC; out: pz, dist, inside, valid
void hm_face_eval(const float* v, float px, float py, float* out) {
    MhFace f; mh_face_setup(v, v + 3, v + 6, 1.0f, &f);
    MhFrag fr; bool ok = mh_face_eval(f, px, py, &fr);
    out[0] = fr.pz; out[1] = fr.dist; out[2] = fr.inside ? 1.f : 0.f; out[3] = (ok && !f.skip) ? 1.f : 0.f;
}
void hm_face_bwd(const float* v, float px, float py, float gz, float gd, float* g) {
    MhFace f; mh_face_setup(v, v + 3, v + 6, 1.0f, &f);
    for (int i = 0; i < 9; ++i) g[i] = 0.f;
    mh_face_bwd(f, px, py, gz, gd, g);
}
}
