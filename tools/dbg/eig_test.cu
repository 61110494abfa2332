#include <cstdio>
#include "../../minicv_b200/csrc/cvmath.cuh"
using namespace mcv::cvm;
__global__ void k(const double* A, double* w, double* V, double* Aout) {
    double L[81], ww[9], VV[81];
    for (int i = 0; i < 81; ++i) L[i] = A[i];
    jacobi_eigen_sym<9>(L, ww, VV);
    for (int i = 0; i < 9; ++i) w[i] = ww[i];
    for (int i = 0; i < 81; ++i) { V[i] = VV[i]; Aout[i] = L[i]; }
}
int main() {
    double A[81], B[20][9];
    unsigned s = 1;
    for (int i = 0; i < 20; ++i) for (int j = 0; j < 9; ++j) { s = s * 1664525u + 1013904223u; B[i][j] = (double)(s >> 8) / (1 << 24) - 0.5; }
    for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) { double t = 0; for (int k2 = 0; k2 < 20; ++k2) t += B[k2][i] * B[k2][j]; A[i * 9 + j] = t; }
    double hw[9], hV[81], hA[81];
    for (int i = 0; i < 81; ++i) hA[i] = A[i];
    jacobi_eigen_sym<9>(hA, hw, hV);
    double *dA, *dw, *dV, *dAo;
    cudaMalloc(&dA, 81 * 8); cudaMalloc(&dw, 72); cudaMalloc(&dV, 81 * 8); cudaMalloc(&dAo, 81 * 8);
    cudaMemcpy(dA, A, 81 * 8, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(dA, dw, dV, dAo);
    double gw[9], gV[81];
    cudaError_t e = cudaDeviceSynchronize();
    printf("err=%s\n", cudaGetErrorString(e));
    cudaMemcpy(gw, dw, 72, cudaMemcpyDeviceToHost); cudaMemcpy(gV, dV, 81 * 8, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 9; ++i) printf("host w=%.12g dev w=%.12g  V0 host %.6g dev %.6g\n", hw[i], gw[i], hV[i], gV[i]);
    return 0;
}
