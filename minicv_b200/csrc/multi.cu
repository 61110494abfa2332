// multi.cu -- multi-GPU forms of the batched exports (SURVEY.md section 8b "optional int deviceMask", section 8e).
//
// Image pairs are independent units: the pair list is cut into contiguous blocks of equal work, one block per
// device of the mask; one host thread per device drives that device's own context (stream set, workspaces, pinned
// ring) through the single-device export, and every thread writes its disjoint slice of the caller's output
// arrays.  No collective, no shared device state; the call returns when every device has finished.  A caller on the
// reference's side of the boundary (one P/Invoke) reaches every GPU of the box this way.
#include "common.cuh"

#include <algorithm>
#include <thread>
#include <vector>
#include <string>

using namespace mcv;

namespace {

int devices_of(uint32_t mask, std::vector<int>& out) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device; libMiniCVNative has no CPU fallback");
        return -1;
    }
    if (mask == 0) mask = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
    for (int d = 0; d < 32; ++d)
        if (mask & (1u << d)) {
            if (d >= n) { set_error("deviceMask names device %d but only %d are visible", d, n); return -1; }
            out.push_back(d);
        }
    return 0;
}

// contiguous blocks of (nearly) equal total weight; bounds has nDev + 1 entries
void partition(const std::vector<double>& w, int nDev, std::vector<int>& bounds) {
    const int n = (int)w.size();
    double total = 0;
    for (double x : w) total += x;
    bounds.assign(nDev + 1, n);
    bounds[0] = 0;
    double acc = 0;
    int p = 0;
    for (int d = 1; d < nDev; ++d) {
        const double target = total * d / nDev;
        while (p < n && acc + w[p] * 0.5 < target) acc += w[p++];
        bounds[d] = p;
    }
}

// runs fn(slot) on one thread per device slot; the first failure's error string becomes the caller's
template <class F>
int run_slots(int nSlots, F fn, std::vector<int>& rc) {
    rc.assign(nSlots, 0);
    std::vector<std::string> errs(nSlots);
    std::vector<std::thread> th;
    th.reserve(nSlots);
    auto body = [&](int s) {
        clear_error();
        rc[s] = fn(s);
        if (rc[s] < 0) errs[s] = get_error();
    };
    for (int s = 1; s < nSlots; ++s) th.emplace_back(body, s);
    body(0);
    for (auto& t : th) t.join();
    for (int s = 0; s < nSlots; ++s)
        if (rc[s] < 0) { set_error("device slot %d: %s", s, errs[s].c_str()); return -1; }
    return 0;
}

std::vector<int64_t> rebase(const int64_t* off, int p0, int p1) {
    std::vector<int64_t> o((size_t)(p1 - p0) + 1);
    for (int p = p0; p <= p1; ++p) o[p - p0] = off[p] - off[p0];
    return o;
}

int check_offsets(const int64_t* off, int nPairs, const char* who) {
    if (nPairs < 0 || !off) { set_error("%s: bad arguments", who); return -1; }
    if (off[0] != 0) { set_error("%s: offsets must start at 0", who); return -1; }
    for (int p = 0; p < nPairs; ++p)
        if (off[p + 1] < off[p]) { set_error("%s: bad offsets at pair %d", who, p); return -1; }
    return 0;
}

}  // namespace

extern "C" {

MCV_API int cvMatchKnn2BatchMulti(const uint8_t* query, const int64_t* queryOffsets, const uint8_t* train,
                                  const int64_t* trainOffsets, int nPairs, int descBytes, int normType, float ratio,
                                  int* trainIdx, float* dist, uint8_t* ratioPass, uint32_t deviceMask) {
    MCV_API_ENTER();
    if (check_offsets(queryOffsets, nPairs, "cvMatchKnn2BatchMulti") || check_offsets(trainOffsets, nPairs, "cvMatchKnn2BatchMulti"))
        return -1;
    if (nPairs == 0) return 0;
    std::vector<int> devs;
    if (devices_of(deviceMask, devs)) return -1;
    const int nDev = (int)std::min<size_t>(devs.size(), (size_t)nPairs);
    std::vector<double> w((size_t)nPairs);
    for (int p = 0; p < nPairs; ++p)
        w[p] = (double)(queryOffsets[p + 1] - queryOffsets[p]) * (double)(trainOffsets[p + 1] - trainOffsets[p]) + 1.0;
    std::vector<int> bounds;
    partition(w, nDev, bounds);
    const size_t rb = (size_t)descBytes;
    std::vector<int> rc;
    if (run_slots(nDev, [&](int s) -> int {
            const int p0 = bounds[s], p1 = bounds[s + 1];
            if (p1 <= p0) return 0;
            const std::vector<int64_t> qo = rebase(queryOffsets, p0, p1), to = rebase(trainOffsets, p0, p1);
            const int64_t q0 = queryOffsets[p0], t0 = trainOffsets[p0];
            return cvMatchKnn2Batch(query ? query + q0 * rb : nullptr, qo.data(), train ? train + t0 * rb : nullptr, to.data(),
                                    p1 - p0, descBytes, normType, ratio, trainIdx ? trainIdx + 2 * q0 : nullptr,
                                    dist ? dist + 2 * q0 : nullptr, ratioPass ? ratioPass + q0 : nullptr, devs[s]);
        }, rc))
        return -1;
    int total = 0;
    for (int v : rc) total += v;
    return total;
}

MCV_API int cvRansacBatchMulti(const McvRansacParams* prm, const McvPoint2d* a, const McvPoint2d* b,
                               const int64_t* offsets, int nPairs, McvMatx33d* models, uint8_t* masks,
                               int32_t* inlierCounts, uint8_t* ok, int32_t* itersRun, uint32_t deviceMask) {
    MCV_API_ENTER();
    if (!prm || check_offsets(offsets, nPairs, "cvRansacBatchMulti")) return -1;
    if (nPairs == 0) return 0;
    if (!models || !ok) { set_error("cvRansacBatchMulti: null outputs"); return -1; }
    std::vector<int> devs;
    if (devices_of(deviceMask, devs)) return -1;
    const int nDev = (int)std::min<size_t>(devs.size(), (size_t)nPairs);
    std::vector<double> w((size_t)nPairs);
    for (int p = 0; p < nPairs; ++p) w[p] = (double)(offsets[p + 1] - offsets[p]) + 64.0;
    std::vector<int> bounds;
    partition(w, nDev, bounds);
    std::vector<int> rc;
    if (run_slots(nDev, [&](int s) -> int {
            const int p0 = bounds[s], p1 = bounds[s + 1];
            if (p1 <= p0) return 0;
            const std::vector<int64_t> o = rebase(offsets, p0, p1);
            const int64_t r0 = offsets[p0];
            McvRansacParams q = *prm;
            q.pairIndexBase = prm->pairIndexBase + p0;     // counter-based sampling: streams keyed by the global pair index
            return cvRansacBatch(&q, a ? a + r0 : nullptr, b ? b + r0 : nullptr, o.data(), p1 - p0, nullptr, nullptr,
                                 models + p0, masks ? masks + r0 : nullptr, inlierCounts ? inlierCounts + p0 : nullptr,
                                 ok + p0, itersRun ? itersRun + p0 : nullptr, nullptr, nullptr, nullptr, nullptr, devs[s]);
        }, rc))
        return -1;
    int total = 0;
    for (int v : rc) total += v;
    return total;
}

MCV_API int cvRecoverPoseBatchMulti(const RecoverPoseConfig* config, const McvPoint2d* a, const McvPoint2d* b,
                                    const int64_t* offsets, int nPairs, int sampleMode, int maxIters, McvMatx33d* R,
                                    McvVec3d* t, uint8_t* ms, int32_t* good, uint8_t* ok, uint32_t deviceMask) {
    MCV_API_ENTER();
    if (!config || check_offsets(offsets, nPairs, "cvRecoverPoseBatchMulti")) return -1;
    if (nPairs == 0) return 0;
    if (!R || !t || !ok) { set_error("cvRecoverPoseBatchMulti: null outputs"); return -1; }
    std::vector<int> devs;
    if (devices_of(deviceMask, devs)) return -1;
    const int nDev = (int)std::min<size_t>(devs.size(), (size_t)nPairs);
    std::vector<double> w((size_t)nPairs);
    for (int p = 0; p < nPairs; ++p) w[p] = (double)(offsets[p + 1] - offsets[p]) + 64.0;
    std::vector<int> bounds;
    partition(w, nDev, bounds);
    std::vector<int> rc;
    if (run_slots(nDev, [&](int s) -> int {
            const int p0 = bounds[s], p1 = bounds[s + 1];
            if (p1 <= p0) return 0;
            const std::vector<int64_t> o = rebase(offsets, p0, p1);
            const int64_t r0 = offsets[p0];
            return cvRecoverPoseBatch(config, a ? a + r0 : nullptr, b ? b + r0 : nullptr, o.data(), p1 - p0, sampleMode,
                                      maxIters, R + p0, t + p0, ms ? ms + r0 : nullptr, good ? good + p0 : nullptr,
                                      ok + p0, devs[s]);
        }, rc))
        return -1;
    int total = 0;
    for (int v : rc) total += v;
    return total;
}

MCV_API int cvSolvePnPRansacBatchMulti(const McvPnpParams* prm, const McvPoint3d* obj, const McvPoint2d* img,
                                       const int64_t* offsets, int nProblems, const double* cams, McvVec3d* rvecs,
                                       McvVec3d* tvecs, uint8_t* masks, int32_t* inlierCounts, uint8_t* ok,
                                       int32_t* itersRun, uint32_t deviceMask) {
    MCV_API_ENTER();
    if (!prm || check_offsets(offsets, nProblems, "cvSolvePnPRansacBatchMulti")) return -1;
    if (nProblems == 0) return 0;
    if (!rvecs || !tvecs || !ok || !cams) { set_error("cvSolvePnPRansacBatchMulti: null arguments"); return -1; }
    std::vector<int> devs;
    if (devices_of(deviceMask, devs)) return -1;
    const int nDev = (int)std::min<size_t>(devs.size(), (size_t)nProblems);
    std::vector<double> w((size_t)nProblems);
    for (int p = 0; p < nProblems; ++p) w[p] = (double)(offsets[p + 1] - offsets[p]) + 64.0;
    std::vector<int> bounds;
    partition(w, nDev, bounds);
    std::vector<int> rc;
    if (run_slots(nDev, [&](int s) -> int {
            const int p0 = bounds[s], p1 = bounds[s + 1];
            if (p1 <= p0) return 0;
            const std::vector<int64_t> o = rebase(offsets, p0, p1);
            const int64_t r0 = offsets[p0];
            McvPnpParams q = *prm;
            q.problemIndexBase = prm->problemIndexBase + p0;
            return cvSolvePnPRansacBatch(&q, obj ? obj + r0 : nullptr, img ? img + r0 : nullptr, o.data(), p1 - p0,
                                         cams + 8 * (size_t)p0, nullptr, nullptr, rvecs + p0, tvecs + p0,
                                         masks ? masks + r0 : nullptr, inlierCounts ? inlierCounts + p0 : nullptr, ok + p0,
                                         itersRun ? itersRun + p0 : nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                         devs[s]);
        }, rc))
        return -1;
    int total = 0;
    for (int v : rc) total += v;
    return total;
}

}  // extern "C"
