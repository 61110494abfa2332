// Tests.fsx -- assertion-based checks for the GPU path through the F# surface (SURVEY.md section 8f rank 2: the
// reference's src/Test/Program.fs has no assertions at all).  Run from the MiniCV repo with the GPU build of
// libMiniCVNative.so on the native search path:   dotnet fsi integration/fsharp/Tests.fsx
// NOT RUN IN THIS REPOSITORY (no dotnet in the build image); the same checks run in Python:
//   matcher: tests/test_knn_gpu.py    H / F: tests/test_ransac_gpu.py    pose: tests/test_ransac_gpu.py    scale: tests/test_camera_pose_gpu.py
#r "nuget: Aardvark.Base, 5.3.0"
#r "bin/Release/net8.0/MiniCV.dll"

open System
open Aardvark.Base
open MiniCV

let mutable failed = 0
let check name cond =
    if cond then printfn "ok   %s" name
    else
        failed <- failed + 1
        printfn "FAIL %s" name

let rnd = Random(1)

// ---- matcher: exact Hamming distances, lowest-index tie-break, ratio flag -----------------------------------------
let popcnt (x : byte) = let mutable v = int x in let mutable c = 0 in (while v <> 0 do c <- c + (v &&& 1); v <- v >>> 1); c
let hamming (a : byte[]) (ao : int) (b : byte[]) (bo : int) n = Seq.init n (fun i -> popcnt (a.[ao + i] ^^^ b.[bo + i])) |> Seq.sum
do
    let nq, nt, db = 200, 300, 32
    let train = Array.init (nt * db) (fun _ -> byte (rnd.Next 256))
    let query = Array.init (nq * db) (fun _ -> byte (rnd.Next 256))
    Array.blit train (7 * db) query (3 * db) db                 // query 3 == train 7
    Array.blit train (7 * db) train (19 * db) db                // ... and train 19 is a duplicate: the tie must go to 7
    let mk (d : byte[]) n = { points = Array.zeroCreate n; descriptors = d :> Array; descriptorDimension = db } : ImageFeatures
    let cnt, idx, dist, pass = OpenCVGpu.matchKnn2 0.8f (mk query nq) (mk train nt)
    let bruteforce i =
        [| for j in 0 .. nt - 1 -> hamming query (i * db) train (j * db) db, j |] |> Array.sortBy id |> Array.take 2
    let allEqual =
        Seq.forall (fun i ->
            let bf = bruteforce i
            idx.[2 * i] = snd bf.[0] && idx.[2 * i + 1] = snd bf.[1] &&
            dist.[2 * i] = float32 (fst bf.[0]) && dist.[2 * i + 1] = float32 (fst bf.[1]) &&
            (pass.[i] <> 0uy) = (dist.[2 * i] < 0.8f * dist.[2 * i + 1])) (seq { 0 .. nq - 1 })
    check "matchKnn2 == brute force (indices, distances, ratio flags)" allEqual
    check "matchKnn2 tie -> lowest train index" (idx.[6] = 7 && idx.[7] = 19 && dist.[6] = 0.0f)
    check "matchKnn2 count" (cnt = (pass |> Array.sumBy (fun b -> if b <> 0uy then 1 else 0)))

// ---- homography: planted model recovered, mask = inliers, deterministic -------------------------------------------
do
    let Hgt = M33d(1.02, 0.08, 12.0, -0.09, 1.03, -7.0, -1.6e-4, -0.7e-4, 1.0)
    let n = 2000
    let a = Array.init n (fun _ -> V2d(rnd.NextDouble() * 1000.0, rnd.NextDouble() * 1000.0))
    let b = a |> Array.mapi (fun i p ->
                if i % 10 < 3 then (let q = Hgt * V3d(p.X, p.Y, 1.0) in V2d(q.X / q.Z, q.Y / q.Z))
                else V2d(rnd.NextDouble() * 1000.0, rnd.NextDouble() * 1000.0))
    match OpenCVGpu.findHomography 3.0 a b, OpenCVGpu.findHomography 3.0 a b with
    | Some (H, mask), Some (H2, mask2) ->
        check "findHomography recovers the planted H" (Fun.ApproximateEquals(H, Hgt, 1e-6))
        check "findHomography mask marks the planted inliers" (Seq.forall (fun i -> (mask.[i] <> 0uy) || i % 10 >= 3) (seq { 0 .. n - 1 }))
        check "findHomography is deterministic" (H = H2 && mask = mask2)
    | _ -> check "findHomography succeeded" false
    check "findHomography fails on 3 points" ((OpenCVGpu.findHomography 3.0 (Array.take 3 a) (Array.take 3 b)).IsNone)

// ---- recoverPose: sentinels untouched on failure (OpenCV.fs:857,866) -------------------------------------------------
do
    let cfg = RecoverPoseConfig.Default
    let res, m, t, ms = OpenCV.recoverPose cfg [| V2d.OO; V2d.IO |] [| V2d.OO; V2d.IO |]
    check "recoverPose: N < 5 -> 0 and the sentinel t" (res = 0 && t = V3d(100.0, 123.0, 432.0) && m = M33d.Identity)

printfn "%s" (if failed = 0 then "ALL PASSED" else sprintf "%d FAILED" failed)
exit failed
