// staging.cuh -- host-buffer plumbing shared by the host-pointer exports (staging.cu).
#pragma once
#include "common.cuh"

namespace mcv {

// true when `p` is ordinary pageable host memory (neither cudaMallocHost / cudaHostAlloc nor cudaHostRegister-ed)
bool host_is_pageable(const void* p);
// Host -> device copy of `bytes` on stream `s`.  Pinned sources: one cudaMemcpyAsync.  Pageable sources of 256 KiB
// and more: through the context's ring of pinned slots, filled by the copy-thread pool; returns once every piece has
// been read from `hSrc` and its DMA is queued on `s` (stream order as for cudaMemcpyAsync).  0 or -1.
int h2d_copy(Context* c, void* dDst, const void* hSrc, size_t bytes, cudaStream_t s, bool pageable);
// Several copies staged as ONE batch of pieces, so that all copy threads stay busy when the individual copies are
// only a few slots long (the query and the train rows of one chunk of pairs).
struct H2dSeg { void* dst; const void* src; size_t bytes; bool pageable; };
int h2d_copy_segs(Context* c, const H2dSeg* segs, int nSegs, cudaStream_t s);
// memcpy split over the copy-thread pool (large copies only)
void parallel_memcpy(void* dst, const void* src, size_t bytes);
int copy_threads();

}  // namespace mcv
