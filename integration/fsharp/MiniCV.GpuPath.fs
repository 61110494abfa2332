// MiniCV.GpuPath.fs -- the F# surface for the exports libMiniCVNative.so adds on the match / estimate path
// (SURVEY.md section 8f rank 2).  Drop next to src/MiniCV/OpenCV.fs (after it in MiniCV.fsproj): it only ADDS
// members; every existing DllImport and wrapper of OpenCV.fs stays as it is (src/MiniCV/OpenCV.fs:339-382, 855-920).
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no dotnet / fsharpc.  The executable stand-in with the same
// names, argument order and sentinels is the Python mirror (minicv_b200/matching.py, estimation.py, camera_pose.py),
// which the parity tests drive through the very same C ABI (include/MiniCVNative.h).
namespace MiniCV

open System
open System.Runtime.InteropServices
open System.Security
open Aardvark.Base

#nowarn "9"
#nowarn "51"

/// McvRansacParams (include/MiniCVNative.h), 72 bytes, sequential.
[<Struct; StructLayout(LayoutKind.Sequential)>]
type RansacParams =
    val mutable model       : int       // 0 = homography, 1 = fundamental, 2 = essential
    val mutable sampleMode  : int       // 0 = OpenCV's RNG stream, 1 = counter-based, 2 = injected
    val mutable stopMode    : int       // 0 = adaptive (OpenCV), 1 = fixed hypothesis count
    val mutable maxIters    : int
    val mutable threshold   : float
    val mutable confidence  : float
    val mutable seed        : uint64
    val mutable focal       : float
    val mutable ppx         : float
    val mutable ppy         : float
    val mutable refine      : int
    val mutable reserved    : int

/// McvCamera: blittable mirror of the Camera record (Camera.fs:7-14), 112 bytes.
[<Struct; StructLayout(LayoutKind.Sequential)>]
type NativeCamera =
    val mutable location : V3d
    val mutable forward  : V3d
    val mutable up       : V3d
    val mutable right    : V3d
    val mutable focal    : V2d
    new(c : Camera) = { location = c.location; forward = c.forward; up = c.up; right = c.right; focal = c.focal }

module GpuNative =
    [<Literal>]
    let lib = "MiniCVNative"

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern nativeint mcvLastError()

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern int mcvDeviceCount()

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern int cvMatchKnn2(byte[] query, int nQuery, byte[] train, int nTrain, int descBytes, int normType,
                           float32 ratio, int[] trainIdx, float32[] dist, byte[] ratioPass)

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern int cvMatchKnn2Batch(byte[] query, int64[] queryOffsets, byte[] train, int64[] trainOffsets, int nPairs,
                                int descBytes, int normType, float32 ratio, int[] trainIdx, float32[] dist,
                                byte[] ratioPass, int device)

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern int cvFindHomography(V2d[] a, V2d[] b, int N, float thr, int maxIters, float conf, M33d& H, byte[] mask)

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern int cvFindFundamentalMat(V2d[] a, V2d[] b, int N, float thr, float conf, int maxIters, M33d& F, byte[] mask)

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern int cvRansacBatch(RansacParams& prm, V2d[] a, V2d[] b, int64[] offsets, int nPairs, int[] injSamples,
                             float[] injModels, M33d[] models, byte[] masks, int[] inlierCounts, byte[] ok,
                             int[] itersRun, int[] best, int[] dbgCounts, float[] dbgModels, int[] dbgSamples, int device)

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern int cvRecoverPoseBatch(RecoverPoseConfig& cfg, V2d[] a, V2d[] b, int64[] offsets, int nPairs, int sampleMode,
                                  int maxIters, M33d[] R, V3d[] t, byte[] ms, int[] good, byte[] ok, int device)

    [<DllImport(lib); SuppressUnmanagedCodeSecurity>]
    extern float cvFindScaled(float inlierThreshold, NativeCamera& srcCam, V3d[] world, V2d[] obs, int n,
                              CameraPose& pose, CameraPose& scaled)

/// Wrappers in the style of module OpenCV (curried arguments, tuples back; OpenCV.fs:855-920).
module OpenCVGpu =
    let private lastError () = Marshal.PtrToStringAnsi(GpuNative.mcvLastError())

    /// kNN match k = 2 + Lowe ratio over binary descriptors (ORB 32 B, AKAZE 61 B, BRISK 64 B): for query i,
    /// trainIdx.[2i..2i+1] / dist.[2i..2i+1] are cv2.BFMatcher(NORM_HAMMING).knnMatch's two best (lowest index wins
    /// ties), pass.[i] <> 0uy iff dist.[2i] < ratio * dist.[2i+1].  Returns the number of passing queries first.
    let matchKnn2 (ratio : float32) (q : ImageFeatures) (t : ImageFeatures) =
        let qd = q.descriptors :?> byte[]
        let td = t.descriptors :?> byte[]
        let n = q.points.Length
        let idx : int[] = Array.zeroCreate (2 * n)
        let dist : float32[] = Array.zeroCreate (2 * n)
        let pass : byte[] = Array.zeroCreate n
        let cnt = GpuNative.cvMatchKnn2(qd, n, td, t.points.Length, q.descriptorDimension, 6, ratio, idx, dist, pass)
        if cnt < 0 then failwithf "cvMatchKnn2: %s" (lastError ())
        cnt, idx, dist, pass

    /// == cv2.findHomography(a, b, RANSAC, thr, maxIters = 2000, confidence = 0.995).  None on failure.
    let findHomography (thr : float) (a : V2d[]) (b : V2d[]) =
        let mutable H = M33d.Identity
        let mask : byte[] = Array.zeroCreate a.Length
        if GpuNative.cvFindHomography(a, b, a.Length, thr, 2000, 0.995, &H, mask) <> 0 then Some (H, mask) else None

    /// == cv2.findFundamentalMat(a, b, FM_RANSAC, thr, 0.99, maxIters = 1000).  None on failure.
    let findFundamentalMat (thr : float) (a : V2d[]) (b : V2d[]) =
        let mutable F = M33d.Identity
        let mask : byte[] = Array.zeroCreate a.Length
        if GpuNative.cvFindFundamentalMat(a, b, a.Length, thr, 0.99, 1000, &F, mask) <> 0 then Some (F, mask) else None

    /// recoverPose for many image pairs in one call: pair p owns rows [offsets.[p], offsets.[p+1]) of a / b.
    let recoverPoseBatch (cfg : RecoverPoseConfig) (a : V2d[]) (b : V2d[]) (offsets : int64[]) =
        let n = offsets.Length - 1
        let R : M33d[] = Array.zeroCreate n
        let t : V3d[] = Array.zeroCreate n
        let ms : byte[] = Array.zeroCreate a.Length
        let good : int[] = Array.zeroCreate n
        let ok : byte[] = Array.zeroCreate n
        let mutable c = cfg
        if GpuNative.cvRecoverPoseBatch(&c, a, b, offsets, n, 0, 1000, R, t, ms, good, ok, -1) < 0 then
            failwithf "cvRecoverPoseBatch: %s" (lastError ())
        R, t, ms, good, ok

module CameraPoseGpu =
    /// Drop-in for CameraPose.findScaled (CameraPose.fs:39-134): same arguments, same (cost, pose) result.
    let findScaledGpu (inlierThreshold : float) (srcCam : Camera) (worldObservations : list<V3d * V2d>) (pose : CameraPose) =
        let world, obs = worldObservations |> List.toArray |> Array.unzip
        let mutable cam = NativeCamera(srcCam)
        let mutable p = pose
        let mutable res = CameraPose()
        let cost = GpuNative.cvFindScaled(inlierThreshold, &cam, world, obs, world.Length, &p, &res)
        cost, res
