// hostmath.cpp -- TEST-ONLY host build of minicv_b200/csrc/cvmath.cuh (the same source the
// CUDA kernels include), so the solver/scoring math can be checked against the oracle in
// the CPU test-suite.  Never linked into libMiniCVNative.so; not a fallback path.
#include "../../minicv_b200/csrc/cvmath.cuh"
#include <string.h>

using namespace mcv::cvm;

extern "C" {

int hm_h_solve4(const float* src, const float* dst, double* H) { return h_solve4(src, dst, H) ? 1 : 0; }
int hm_h_check_subset(const float* src, const float* dst) { return h_check_subset(src, dst) ? 1 : 0; }
int hm_have_collinear(const float* pts, int count) { return have_collinear(pts, count) ? 1 : 0; }
void hm_h_error(const double* H, const float* a, const float* b, int n, float* err) {
    float Hf[8];
    for (int i = 0; i < 8; ++i) Hf[i] = (float)H[i];
    for (int i = 0; i < n; ++i) err[i] = h_error(Hf, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}
int hm_f_solve7(const float* a, const float* b, double* F) { return f_solve7(a, b, F); }
void hm_f_error(const double* F, const float* a, const float* b, int n, float* err) {
    for (int i = 0; i < n; ++i) err[i] = f_error(F, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}
int hm_e_solve5(const double* a, const double* b, double* Es) { return e_solve5(a, b, Es); }
void hm_e_error(const double* E, const double* a, const double* b, int n, float* err) {
    for (int i = 0; i < n; ++i) err[i] = e_error(E, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}
void hm_f_inlier(const double* F, const float* a, const float* b, int n, float thr, unsigned char* out) {
    for (int i = 0; i < n; ++i) out[i] = f_inlier(F, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], thr) ? 1 : 0;
}
void hm_e_inlier(const double* E, const double* a, const double* b, int n, float thr, unsigned char* out) {
    for (int i = 0; i < n; ++i) out[i] = e_inlier(E, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], thr) ? 1 : 0;
}
void hm_decompose_essential(const double* E, double* R1, double* R2, double* t) { decompose_essential(E, R1, R2, t); }
void hm_cheirality(const double* R, const double* t, const double* a, const double* b, int n, double dist,
                   unsigned char* ok) {
    for (int i = 0; i < n; ++i) ok[i] = cheirality_ok(R, t, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], dist) ? 1 : 0;
}
void hm_svd3(const double* A, double* w, double* u, double* vt) { svd_square<3>(A, w, u, vt); }
void hm_svd4(const double* A, double* w, double* u, double* vt) { svd_square<4>(A, w, u, vt); }
void hm_solve_sym_eig8(const double* A, const double* b, double* x) { solve_sym_eig<8>(A, b, x); }
void hm_solve_sym_chol8(const double* A, const double* b, double* x) { solve_sym_chol_or_eig<8>(A, b, x); }
void hm_eigen_sym9(const double* A, double* w, double* V) {
    double M[81]; memcpy(M, A, sizeof(M));
    jacobi_eigen_sym<9>(M, w, V);
}
int hm_update_num_iters(double p, double ep, int mp, int maxIters) { return update_num_iters(p, ep, mp, maxIters); }
void hm_mwc(int n, unsigned* out) { Mwc r; for (int i = 0; i < n; ++i) out[i] = r.next(); }
int hm_real_roots10(const double* c, double* out) { return real_roots10(c, out); }

}

// ---- PnP row (pnpmath.cuh) ------------------------------------------------------------------------
#include "../../minicv_b200/csrc/pnpmath.cuh"
extern "C" {
// the reference's solveAp3p (ap3p.cpp:282): img [3][2], obj [3][3]
int hm_ap3p(const double* img, const double* obj, double inv_fx, double inv_fy, double cx_fx, double cy_fy, double* Rs,
            double* ts) {
    double b[3][3];
    for (int i = 0; i < 3; ++i) ap3p_bearing(img[2 * i], img[2 * i + 1], inv_fx, inv_fy, cx_fx, cy_fy, b[i], nullptr);
    return ap3p_poses(b[0], b[1], b[2], obj, obj + 3, obj + 6, false, nullptr, nullptr, Rs, ts);
}
int hm_pnp_solve4(const double* cam, const float* obj, const float* img, double* rvec, double* tvec) {
    PnpCamera c{cam[0], cam[1], cam[2], cam[3], cam[4], cam[5], cam[6], cam[7]};
    return pnp_solve4(c, obj, img, rvec, tvec) ? 1 : 0;
}
void hm_rodrigues_v2m(const double* r, double* R) { rodrigues_v2m(r, R); }
void hm_rodrigues_m2v(const double* R, double* r) { rodrigues_m2v(R, r); }
void hm_undistort(const double* cam, const float* uv, int n, float* out) {
    PnpCamera c{cam[0], cam[1], cam[2], cam[3], cam[4], cam[5], cam[6], cam[7]};
    for (int i = 0; i < n; ++i) {
        double x, y;
        undistort_point(c, (double)uv[2 * i], (double)uv[2 * i + 1], &x, &y);
        out[2 * i] = (float)x; out[2 * i + 1] = (float)y;
    }
}
void hm_project_error(const double* cam, const double* rvec, const double* tvec, const float* obj, const float* img, int n,
                      float* err) {
    PnpCamera c{cam[0], cam[1], cam[2], cam[3], cam[4], cam[5], cam[6], cam[7]};
    double R[9];
    rodrigues_v2m(rvec, R);
    for (int i = 0; i < n; ++i)
        err[i] = project_error(R, tvec, c, c.distorted(), obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i], img[2 * i + 1]);
}
// cv::solvePnP(SOLVEPNP_EPNP) on float64 points: undistort, EPnP, Rodrigues
void hm_epnp(const double* cam, const double* obj, const double* img, int n, double* rvec, double* tvec) {
    PnpCamera c{cam[0], cam[1], cam[2], cam[3], cam[4], cam[5], cam[6], cam[7]};
    double* und = new double[2 * n];
    for (int i = 0; i < n; ++i) undistort_point(c, img[2 * i], img[2 * i + 1], und + 2 * i, und + 2 * i + 1);
    double R[9];
    epnp_sequential(n, obj, und, c.fx, c.fy, c.cx, c.cy, R, tvec);
    rodrigues_m2v(R, rvec);
    delete[] und;
}
}

// ---- scale-search row (scalemath.cuh) ---------------------------------------------------------------
#include "../../minicv_b200/csrc/scalemath.cuh"

extern "C" {
using namespace mcv::scm;

static Cam cam_from_row(const double* c) {
    Cam cam;
    for (int i = 0; i < 3; ++i) { cam.loc[i] = c[i]; cam.fwd[i] = c[3 + i]; cam.up[i] = c[6 + i]; cam.right[i] = c[9 + i]; }
    cam.focal[0] = c[12]; cam.focal[1] = c[13];
    return cam;
}
int hm_lu_inverse4(const double* A, double* inv) { return lu_inverse4(A, inv) ? 1 : 0; }
// camera (14 doubles, McvCamera order) that scale s produces
void hm_camera_for_scale(const double* camRow, const double* R, const double* T, double s, double* outRow) {
    const Cam src = cam_from_row(camRow);
    double basis[16];
    camera_basis(src, basis);
    Cam out;
    camera_for_scale(basis, src.focal, R, T, s, out);
    for (int i = 0; i < 3; ++i) { outRow[i] = out.loc[i]; outRow[3 + i] = out.fwd[i]; outRow[6 + i] = out.up[i]; outRow[9 + i] = out.right[i]; }
    outRow[12] = out.focal[0]; outRow[13] = out.focal[1];
}
// candidates [n][2], ok [n]
void hm_candidate_scales(const double* camRow, const double* R, const double* T, const double* world, const double* obs,
                         int n, double* cand, unsigned char* ok) {
    ScaleSetup S;
    scale_setup(cam_from_row(camRow), R, T, S);
    for (int i = 0; i < n; ++i) {
        double sx = 0, sy = 0;
        ok[i] = candidate_scales(S, world + 3 * i, obs + 2 * i, sx, sy) ? 1 : 0;
        cand[2 * i] = sx; cand[2 * i + 1] = sy;
    }
}
// avgReprojectionError for each scale
void hm_scale_costs(const double* camRow, const double* R, const double* T, const double* world, const double* obs, int n,
                    const double* scales, int ns, double* costs) {
    const Cam src = cam_from_row(camRow);
    double basis[16];
    camera_basis(src, basis);
    for (int k = 0; k < ns; ++k) {
        Cam cam;
        camera_for_scale(basis, src.focal, R, T, scales[k], cam);
        double sum = 0.0;
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            double term;
            if (project_term(cam, world[3 * i], world[3 * i + 1], world[3 * i + 2], obs[2 * i], obs[2 * i + 1], term)) {
                sum = dadd(sum, term);
                ++cnt;
            }
        }
        costs[k] = cnt ? ddiv(sum, (double)cnt) : INFINITY;
    }
}
}

// ---- fp32 filter in front of f_inlier (cvmath.cuh) ---------------------------------------------------
extern "C" {
// For n (model, point) pairs: filt[i] = filter says "certainly outlier", exact[i] = f_inlier.
void hm_f_filter(const double* F, int nModels, const float* a, const float* b, int n, float thr, unsigned char* filt,
                 unsigned char* exact) {
    const int TILE = 1024;                                       // as in score_kernel: bounds per tile of correspondences
    for (int k = 0; k < nModels; ++k) {
        FFilter ff;
        f_filter_init(F + 9 * k, ff);
        const float thrM = thr * FILTER_MARGIN;
        for (int i0 = 0; i0 < n; i0 += TILE) {
            const int i1 = i0 + TILE < n ? i0 + TILE : n;
            float mx[4] = {0, 0, 0, 0};
            for (int i = i0; i < i1; ++i) {
                const float v[4] = {fabsf(a[2 * i]), fabsf(a[2 * i + 1]), fabsf(b[2 * i]), fabsf(b[2 * i + 1])};
                for (int c = 0; c < 4; ++c) if (!(v[c] <= mx[c])) mx[c] = v[c];        // NaN sticks
            }
            FFilterTile t;
            f_filter_tile(ff, mx[0], mx[1], mx[2], mx[3], t);
            for (int i = i0; i < i1; ++i) {
                filt[(size_t)k * n + i] = f_certain_outlier(ff, t, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], thrM) ? 1 : 0;
                exact[(size_t)k * n + i] = f_inlier(F + 9 * k, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], thr) ? 1 : 0;
            }
        }
    }
}
}

extern "C" {
// The same for e_inlier: points are fp64 (normalised coordinates), the filter sees their fp32 roundings.
void hm_e_filter(const double* E, int nModels, const double* a, const double* b, int n, float thr, unsigned char* filt,
                 unsigned char* exact) {
    const int TILE = 512;
    for (int k = 0; k < nModels; ++k) {
        FFilter ff;
        f_filter_init(E + 9 * k, ff);
        const float thrM = thr * FILTER_MARGIN;
        for (int i0 = 0; i0 < n; i0 += TILE) {
            const int i1 = i0 + TILE < n ? i0 + TILE : n;
            float mx[4] = {0, 0, 0, 0};
            for (int i = i0; i < i1; ++i) {
                const float v[4] = {fabsf((float)a[2 * i]), fabsf((float)a[2 * i + 1]), fabsf((float)b[2 * i]), fabsf((float)b[2 * i + 1])};
                for (int c = 0; c < 4; ++c) if (!(v[c] <= mx[c])) mx[c] = v[c];
            }
            FFilterTile t;
            e_filter_tile(ff, mx[0], mx[1], mx[2], mx[3], t);
            for (int i = i0; i < i1; ++i) {
                filt[(size_t)k * n + i] = e_certain_outlier(ff, t, (float)a[2 * i], (float)a[2 * i + 1], (float)b[2 * i],
                                                            (float)b[2 * i + 1], thrM) ? 1 : 0;
                exact[(size_t)k * n + i] = e_inlier(E + 9 * k, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], thr) ? 1 : 0;
            }
        }
    }
}
}

extern "C" {
// PnP filter vs the exact reprojection test: nModels (rvec, tvec) x n correspondences; cam = fx, fy, cx, cy (no distortion).
void hm_pnp_filter(const double* models, int nModels, const float* obj, const float* img, int n, const double* cam4,
                   float thr, unsigned char* filt, unsigned char* exact) {
    PnpCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    cam.k1 = cam.k2 = cam.p1 = cam.p2 = 0;
    const int TILE = 512;
    for (int k = 0; k < nModels; ++k) {
        double R[9];
        rodrigues_v2m(models + 6 * k, R);
        const double* t = models + 6 * k + 3;
        PnpFilter pf;
        pnp_filter_init(R, t, cam, false, pf);
        for (int i0 = 0; i0 < n; i0 += TILE) {
            const int i1 = i0 + TILE < n ? i0 + TILE : n;
            float mx[7] = {0, 0, 0, 0, 0, 0, 0};
            for (int i = i0; i < i1; ++i) {
                const float v[7] = {obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], (float)((double)img[2 * i] - cam.cx),
                                    (float)((double)img[2 * i + 1] - cam.cy), img[2 * i], img[2 * i + 1]};
                for (int c = 0; c < 7; ++c) if (!(fabsf(v[c]) <= mx[c])) mx[c] = fabsf(v[c]);
            }
            PnpFilterTile ft;
            pnp_filter_tile(pf, mx[0], mx[1], mx[2], mx[3], mx[4], mx[5], mx[6], thr, ft);
            for (int i = i0; i < i1; ++i) {
                const float uc = (float)((double)img[2 * i] - cam.cx), vc = (float)((double)img[2 * i + 1] - cam.cy);
                filt[(size_t)k * n + i] = pnp_certain_outlier(pf, ft, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], uc, vc) ? 1 : 0;
                exact[(size_t)k * n + i] =
                    project_error(R, t, cam, false, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i], img[2 * i + 1]) <= thr ? 1 : 0;
            }
        }
    }
}
}

extern "C" {
// Homography filter vs the exact float32 test.  H: nModels x 9 doubles (h33 = 1 normalisation as the kernels use).
void hm_h_filter(const double* H, int nModels, const float* a, const float* b, int n, float thr, unsigned char* filt,
                 unsigned char* exact) {
    const int TILE = 1024;
    for (int k = 0; k < nModels; ++k) {
        float Hf[8];
        for (int i = 0; i < 8; ++i) Hf[i] = (float)H[9 * k + i];
        for (int i0 = 0; i0 < n; i0 += TILE) {
            const int i1 = i0 + TILE < n ? i0 + TILE : n;
            float mx[4] = {0, 0, 0, 0};
            for (int i = i0; i < i1; ++i) {
                const float v[4] = {fabsf(a[2 * i]), fabsf(a[2 * i + 1]), fabsf(b[2 * i]), fabsf(b[2 * i + 1])};
                for (int c = 0; c < 4; ++c) if (!(v[c] <= mx[c])) mx[c] = v[c];
            }
            HFilterTile t;
            h_filter_tile(Hf, mx[0], mx[1], mx[2], mx[3], thr, t);
            for (int i = i0; i < i1; ++i) {
                filt[(size_t)k * n + i] = h_certain_outlier(Hf, t, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]) ? 1 : 0;
                exact[(size_t)k * n + i] = h_error(Hf, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]) <= thr ? 1 : 0;
            }
        }
    }
}
}

// ---- level-0 (tensor-core) filter of the F / E scoring stage (filter0.cuh) ------------------------------------------
#include "../../minicv_b200/csrc/filter0.cuh"
#include <vector>
#include <algorithm>
#include <numeric>
extern "C" {
// isE: 0 fundamental, 1 essential, 2 homography (M rows = H with h33 = 1; no classifier, cls = exact).
// Emulates score_tc.cu on the host: per-pair normalisation from the finite coordinate ranges, correspondences sorted by
// image-1 cell (sorted != 0) or left in input order (boxes as large as the image: must still be conservative), split-TF32
// rows for models and correspondences, stage boxes and thresholds, r~ accumulated over the 27 split terms in fp32 in index
// order (mode 0), in reverse order (mode 1) or in fp64 (mode 2) -- the tensor core's own accumulation order is unspecified;
// its error is budgeted, these three bracket it.  Outputs in INPUT order:
// filt[k*n+i] = level 0 says "certainly outlier", exact[k*n+i] = the reference's exact test,
// cls[k*n+i] = the two-sided fp32 classifier (0 outlier, 1 inlier, 2 undecided).
void hm_filter0(int isE, const double* M, int nModels, const double* a_in, const double* b_in, int n, float thr, int mode, int sorted,
                unsigned char* filt, unsigned char* exact, unsigned char* cls) {
    const int NT = 64;
    // H and F work on the float32 view of the coordinates (prep_points_kernel), E on the fp64 normalised coordinates
    std::vector<double> av(a_in, a_in + 2 * (size_t)n), bv(b_in, b_in + 2 * (size_t)n);
    if (isE != 1) for (size_t i = 0; i < av.size(); ++i) { av[i] = (double)(float)av[i]; bv[i] = (double)(float)bv[i]; }
    const double* a = av.data();
    const double* b = bv.data();
    double lo[4] = {DBL_MAX, DBL_MAX, DBL_MAX, DBL_MAX}, hi[4] = {-DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
    float mx[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const double c[4] = {a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]};
        for (int j = 0; j < 4; ++j) {
            if (fabs(c[j]) <= DBL_MAX) { if (c[j] < lo[j]) lo[j] = c[j]; if (c[j] > hi[j]) hi[j] = c[j]; }
            const float v = fabsf((float)c[j]);
            if (v != v || (mx[j] == mx[j] && v > mx[j])) mx[j] = v;    // a NaN sticks (the kernel takes the max of bit patterns)
        }
    }
    PairNorm pn;
    pair_norm_init(lo, hi, pn);
    pn.X1 = mx[0]; pn.Y1 = mx[1]; pn.X2 = mx[2]; pn.Y2 = mx[3];
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::vector<Hat> hat(n);
    for (int i = 0; i < n; ++i) hat[i] = hat_coords(pn, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
    if (sorted) std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return hat_cell(hat[x]) < hat_cell(hat[y]); });
    std::vector<float> rows((size_t)n * F0_K);
    double pmax = 0, vmax = 0;
    for (int s = 0; s < n; ++s) {
        double v2;
        const double n2 = isE == 2 ? filter0_features_h(hat[order[s]], rows.data() + (size_t)s * F0_K)
                                   : filter0_features(hat[order[s]], rows.data() + (size_t)s * F0_K, &v2);
        if (n2 > pmax) pmax = n2;
        if (v2 > vmax) vmax = v2;
    }
    pn.psiMax = (float)(sqrt(pmax) * (1.0 + 1e-6)) * 1.0000002f;
    pn.v2Max = (float)(vmax * (1.0 + 1e-6)) * 1.0000002f;
    const int nst = (n + NT - 1) / NT;
    std::vector<StageBox> box(nst);
    for (int st = 0; st < nst; ++st) {
        double ulo = DBL_MAX, uhi = -DBL_MAX, vlo = DBL_MAX, vhi = -DBL_MAX;
        for (int s = st * NT; s < n && s < (st + 1) * NT; ++s) {
            const Hat& h = hat[order[s]];
            ulo = fmin(ulo, h.u1); uhi = fmax(uhi, h.u1); vlo = fmin(vlo, h.v1); vhi = fmax(vhi, h.v1);
        }
        StageBox bx;
        bx.cu = (float)(0.5 * ulo + 0.5 * uhi); bx.cv = (float)(0.5 * vlo + 0.5 * vhi);
        double pu = 0, pv = 0;
        for (int s = st * NT; s < n && s < (st + 1) * NT; ++s) {
            const Hat& h = hat[order[s]];
            pu = fmax(pu, fabs(h.u1 - (double)bx.cu)); pv = fmax(pv, fabs(h.v1 - (double)bx.cv));
        }
        bx.pu = (float)(pu * (1.0 + 1e-6)) * 1.0000002f; bx.pv = (float)(pv * (1.0 + 1e-6)) * 1.0000002f;
        box[st] = bx;
    }
    for (int k = 0; k < nModels; ++k) {
        float fr[F0_K];
        Model0 m0;
        float Hf[8];
        for (int j = 0; j < 8; ++j) Hf[j] = (float)M[9 * k + j];
        if (isE == 2) filter0_model_h(pn, Hf, thr, fr, m0);
        else filter0_model(isE, pn, M + 9 * k, thr, fr, m0);
        const double thrk = (double)thr * pn.k * pn.k;
        const float thrkM = (float)(thrk * (double)CLASS_MARGIN), thrkL = (float)(thrk * (double)CLASS_LOW);
        for (int s = 0; s < n; ++s) {
            const int i = order[s];
            const float T = isE == 2 ? filter0_threshold_h(m0, box[s / NT]) : filter0_threshold(isE, m0, box[s / NT]);
            const float* p = rows.data() + (size_t)s * F0_K;
            float r;
            if (mode == 2) { double acc = 0; for (int j = 0; j < 27; ++j) acc += (double)fr[j] * (double)p[j]; r = (float)acc; }
            else if (mode == 1) { volatile float acc = 0; for (int j = 26; j >= 0; --j) { volatile float t = fr[j] * p[j]; acc = acc + t; } r = acc; }
            else { volatile float acc = 0; for (int j = 0; j < 27; ++j) { volatile float t = fr[j] * p[j]; acc = acc + t; } r = acc; }
            filt[(size_t)k * n + i] = (fabsf(r) > T) ? 1 : 0;
            const Hat& h = hat[i];
            if (isE == 2) {
                const bool in = h_error(Hf, (float)a[2 * i], (float)a[2 * i + 1], (float)b[2 * i], (float)b[2 * i + 1]) <= thr;
                exact[(size_t)k * n + i] = in ? 1 : 0;
                cls[(size_t)k * n + i] = in ? 1 : 0;
                continue;
            }
            cls[(size_t)k * n + i] = (unsigned char)hat_classify(isE, m0, thrkM, thrkL, (float)h.u1, (float)h.v1, (float)h.u2, (float)h.v2);
            if (isE) {
                exact[(size_t)k * n + i] = e_inlier(M + 9 * k, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], thr) ? 1 : 0;
            } else {
                const float x1 = (float)a[2 * i], y1 = (float)a[2 * i + 1], x2 = (float)b[2 * i], y2 = (float)b[2 * i + 1];
                exact[(size_t)k * n + i] = f_inlier(M + 9 * k, (double)x1, (double)y1, (double)x2, (double)y2, thr) ? 1 : 0;
            }
        }
    }
}
}
extern "C" {
void hm_tf32_split(const float* x, int n, float* hi, float* lo) {
    for (int i = 0; i < n; ++i) tf32_split(x[i], hi[i], lo[i]);
}
}
