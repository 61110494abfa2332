// forward.cu -- the reference exports that are NOT on the match / two-view / PnP-RANSAC path
// (plain PnP and its refiners, detectors, fiducials, cvTest: reference src/MiniCVNative/MiniCVNative.cpp:48,93,
// 140,152,221,347,384,439,504 and ap3p.cpp:282).  They are out of scope for acceleration and need
// the OpenCV C++ SDK, so this library does not implement them; it keeps every F# DllImport
// (src/MiniCV/OpenCV.fs:343-382) resolving by forwarding to a side-by-side build of the original
// shim: the library named by $MINICV_ORIGINAL_NATIVE (default "libMiniCVNativeCv.so"), opened
// lazily.  When it is absent each export returns the reference's failure value (false / 0 /
// nullptr / no-op) and records an error string -- never a computed result.
#include "common.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <mutex>

namespace {

void* original_lib() {
    static void* h = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* name = getenv("MINICV_ORIGINAL_NATIVE");
        h = dlopen(name && *name ? name : "libMiniCVNativeCv.so", RTLD_NOW | RTLD_LOCAL);
    });
    return h;
}

template <class Fn> Fn resolve(const char* sym) {
    void* h = original_lib();
    Fn f = h ? reinterpret_cast<Fn>(dlsym(h, sym)) : nullptr;
    if (!f) mcv::set_error("%s is not part of the GPU path and no original MiniCVNative is loaded "
                           "(set MINICV_ORIGINAL_NATIVE)", sym);
    return f;
}


}  // namespace

extern "C" {

MCV_API bool cvSolvePnP(const McvPoint2d* img, const McvPoint3d* world, const int N, const McvMatx33d K,
                        const double* dist, const int solverKind, McvVec3d* tVec, McvVec3d* rVec) {
    using Fn = bool (*)(const McvPoint2d*, const McvPoint3d*, int, McvMatx33d, const double*, int, McvVec3d*, McvVec3d*);
    Fn f = resolve<Fn>("cvSolvePnP");
    return f ? f(img, world, N, K, dist, solverKind, tVec, rVec) : false;
}

}  // extern "C"

// cvSolvePnPRansac is native for solverKind 5 (AP3P) -- api_pnp.cu; every other kind lands here.
namespace mcv {
bool forward_solve_pnp_ransac(const McvPoint2d* img, const McvPoint3d* world, const int N, const McvMatx33d K,
                              const double* dist, const int solverKind, const int iterationsCount,
                              const float reprojectionError, const double confidence, McvVec3d* tVec,
                              McvVec3d* rVec, int* inlierCount, int* outInliers) {
    using Fn = bool (*)(const McvPoint2d*, const McvPoint3d*, int, McvMatx33d, const double*, int, int, float, double,
                        McvVec3d*, McvVec3d*, int*, int*);
    Fn f = resolve<Fn>("cvSolvePnPRansac");
    return f ? f(img, world, N, K, dist, solverKind, iterationsCount, reprojectionError, confidence, tVec, rVec,
                 inlierCount, outInliers)
             : false;
}
}  // namespace mcv

extern "C" {

MCV_API void cvRefinePnPLM(const McvPoint2d* img, const McvPoint3d* world, const int N, const McvMatx33d K,
                           const double* dist, McvVec3d* tVec, McvVec3d* rVec) {
    using Fn = void (*)(const McvPoint2d*, const McvPoint3d*, int, McvMatx33d, const double*, McvVec3d*, McvVec3d*);
    if (Fn f = resolve<Fn>("cvRefinePnPLM")) f(img, world, N, K, dist, tVec, rVec);
}

MCV_API void cvRefinePnPVVS(const McvPoint2d* img, const McvPoint3d* world, const int N, const McvMatx33d K,
                            const double* dist, McvVec3d* tVec, McvVec3d* rVec) {
    using Fn = void (*)(const McvPoint2d*, const McvPoint3d*, int, McvMatx33d, const double*, McvVec3d*, McvVec3d*);
    if (Fn f = resolve<Fn>("cvRefinePnPVVS")) f(img, world, N, K, dist, tVec, rVec);
}

MCV_API McvDetectorResult* cvDetectFeatures(char* data, int width, int height, int channels, int mode, void* config) {
    using Fn = McvDetectorResult* (*)(char*, int, int, int, int, void*);
    Fn f = resolve<Fn>("cvDetectFeatures");
    return f ? f(data, width, height, channels, mode, config) : nullptr;
}

MCV_API void cvFreeFeatures(McvDetectorResult* res) {
    using Fn = void (*)(McvDetectorResult*);
    if (!res) return;
    if (Fn f = resolve<Fn>("cvFreeFeatures")) f(res);
}

MCV_API bool cvDetectQRCode(char* data, int width, int height, int channels, int* positions, int* count) {
    using Fn = bool (*)(char*, int, int, int, int*, int*);
    Fn f = resolve<Fn>("cvDetectQRCode");
    return f ? f(data, width, height, channels, positions, count) : false;
}

MCV_API bool cvDetectArucoMarkers(char* data, int width, int height, int channels, int* infoCount, void* infos) {
    using Fn = bool (*)(char*, int, int, int, int*, void*);
    Fn f = resolve<Fn>("cvDetectArucoMarkers");
    return f ? f(data, width, height, channels, infoCount, infos) : false;
}

MCV_API void cvTest(void) {
    using Fn = void (*)();
    if (Fn f = resolve<Fn>("cvTest")) f();
}

}  // extern "C"
