// In-library launch accounting: every kernel launch goes through LaunchScope, which counts it and,
// when profiling is enabled, brackets it with CUDA events on the launching stream.  bench.py reads
// the per-kernel totals (wtpse_profile_read) to report the dominant kernel's duration measured
// inside the timed region.
#pragma once
#include <cuda_runtime.h>

namespace wtpse {

enum KernelId {
    kKernGram = 0,
    kKernEpilogueFwd,
    kKernEpilogueBwd,
    kKernApply,
    kKernMmdFwd,
    kKernMmdBwd,
    kKernMseFwd,
    kKernMseBwd,
    kKernFuse,
    kKernLabels,
    kKernWaveletFwd,
    kKernWaveletBwd,
    kKernGramReduce,      // counted only: timed inside the kKernEpilogueFwd scope
    kKernMmat,            // counted only: timed inside the kKernApply scope
    kKernUpsample,
    kKernCount
};

const char* kernel_name(int id);
void profile_record_begin(int id, cudaStream_t s);
void profile_record_end(int id, cudaStream_t s);
// a scope that brackets two kernels (chained by programmatic dependent launch, so no event may sit between them)
// counts the second one here
void profile_count_kernel(int id);

struct LaunchScope {
    int id;
    cudaStream_t s;
    LaunchScope(int id_, cudaStream_t s_) : id(id_), s(s_) { profile_record_begin(id, s); }
    ~LaunchScope() { profile_record_end(id, s); }
};

}  // namespace wtpse
