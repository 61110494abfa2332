// l2_order.cuh -- squared L2 distance in OpenCV's accumulation order (hal::normL2Sqr_ for float32, 128-bit
// universal intrinsics, no FMA; see oracle/bf_knn.py l2_matrix_cv, pinned to cv2 4.13.0 by
// tests/golden/knn_l2_general.npz).  Shared by the fp32 engine (l2_knn.cu) and the exact re-scoring pass of the
// tcgen05 engine (l2_tc.cu), so both return cv2's float32 distances bit for bit.
//
// The 16 strided partial sums acc[a][l] = sum_b (x[16 b + 4 a + l])^2 are independent of each other, so instead of
// carrying 16 accumulators per distance they are completed one at a time and folded in the order the SIMD code
// folds them:  s[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l]  for l = 0, 2, 1, 3 and
// d = (s[0] + s[2]) + (s[1] + s[3]); the last D % 16 elements are then added in index order.
#pragma once

namespace mcv {

// One thread, an MT x MT micro-tile.  q / t point at the thread's first row inside tiles stored [k][row] with row
// stride `ld` (element k of row r at [k * ld + r]).  out[i][j] = sum over k of (q_i[k] - t_j[k])^2 in OpenCV's order.
template <int MT>
__device__ __forceinline__ void l2_sqr_cv_order(const float* __restrict__ q, const float* __restrict__ t, int ld,
                                                int dim, float (&out)[MT][MT]) {
    const int nb = dim >> 4;
    float u[MT][MT], v[MT][MT];
#pragma unroll
    for (int li = 0; li < 4; ++li) {
        const int l = (li == 0) ? 0 : (li == 1) ? 2 : (li == 2) ? 1 : 3;
        float s[MT][MT];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float acc[MT][MT];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < MT; ++j) acc[i][j] = 0.f;
            const float* qk = q + (size_t)(4 * a + l) * ld;
            const float* tk = t + (size_t)(4 * a + l) * ld;
#pragma unroll 4
            for (int b = 0; b < nb; ++b) {
                float x[MT], y[MT];
#pragma unroll
                for (int i = 0; i < MT; ++i) { x[i] = qk[i]; y[i] = tk[i]; }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < MT; ++j) {
                        const float df = __fsub_rn(x[i], y[j]);
                        acc[i][j] = __fadd_rn(__fmul_rn(df, df), acc[i][j]);
                    }
                qk += 16 * (size_t)ld;
                tk += 16 * (size_t)ld;
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < MT; ++j) s[i][j] = (a == 0) ? acc[i][j] : __fadd_rn(s[i][j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                if (li == 0) u[i][j] = s[i][j];
                else if (li == 1) u[i][j] = __fadd_rn(u[i][j], s[i][j]);       // s[0] + s[2]
                else if (li == 2) v[i][j] = s[i][j];
                else v[i][j] = __fadd_rn(v[i][j], s[i][j]);                    // s[1] + s[3]
            }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) out[i][j] = __fadd_rn(u[i][j], v[i][j]);
    for (int k = nb << 4; k < dim; ++k) {
        float x[MT], y[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) { x[i] = q[(size_t)k * ld + i]; y[i] = t[(size_t)k * ld + i]; }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const float df = __fsub_rn(x[i], y[j]);
                out[i][j] = __fadd_rn(out[i][j], __fmul_rn(df, df));
            }
    }
}

// Same order for one pair of rows stored contiguously (re-scoring: a.k.a. q[k], t[k]).
__device__ __forceinline__ float l2_sqr_cv_order_rows(const float* __restrict__ q, const float* __restrict__ t, int dim) {
    const int nb = dim >> 4;
    float u = 0.f, v = 0.f;
#pragma unroll
    for (int li = 0; li < 4; ++li) {
        const int l = (li == 0) ? 0 : (li == 1) ? 2 : (li == 2) ? 1 : 3;
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float acc = 0.f;
            for (int b = 0; b < nb; ++b) {
                const int k = 16 * b + 4 * a + l;
                const float df = __fsub_rn(q[k], t[k]);
                acc = __fadd_rn(__fmul_rn(df, df), acc);
            }
            s = (a == 0) ? acc : __fadd_rn(s, acc);
        }
        if (li == 0) u = s;
        else if (li == 1) u = __fadd_rn(u, s);
        else if (li == 2) v = s;
        else v = __fadd_rn(v, s);
    }
    float d = __fadd_rn(u, v);
    for (int k = nb << 4; k < dim; ++k) {
        const float df = __fsub_rn(q[k], t[k]);
        d = __fadd_rn(d, __fmul_rn(df, df));
    }
    return d;
}

}  // namespace mcv
