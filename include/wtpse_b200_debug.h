/*
 * wtpse_b200_debug.h -- diagnostics of libwtpse_b200.so.  NOT part of the product ABI (include/wtpse_b200.h): nothing
 * in the reference-facing path (wt-pse-code_b200/functional.py, dropin.py, segmentation.py) calls these; bench.py, the
 * probe tools and the tests that compare kernel variants do.
 */
#ifndef WTPSE_B200_DEBUG_H
#define WTPSE_B200_DEBUG_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- launch accounting / in-step kernel timing -------------------------------------------------
 * Every kernel launch made by the library is counted; with profiling enabled each launch is also
 * bracketed by CUDA events on its own stream.  Kernel ids: 0 .. wtpse_profile_kernel_count()-1. */
void        wtpse_profile_enable(int on);
void        wtpse_profile_reset(void);
int         wtpse_profile_kernel_count(void);
const char* wtpse_profile_kernel_name(int id);
long long   wtpse_profile_launches(int id);            /* id < 0: all kernels */
/* Sum of event-timed durations (ms) of kernel `id` since the last reset; synchronises those events. */
int         wtpse_profile_read(int id, long long* timed_launches, double* total_ms);

/* ---- variant switches ------------------------------------------------------------------------
 * Process-wide integers read when a call is enqueued (a test that flips one restores it).  The product path never
 * touches them; every value of every switch produces the same results (the tests compare them).
 *   "fused_tail"           1 (default) forward tail inside the Gram kernel (last-arriving CTA); 0 separate kernels
 *   "apply_round_robin"    NCHW apply kernel: tiles dealt round-robin over the CTAs (1, default) or contiguous ranges (0)
 *   "l2_hint"              L2 evict-first policy on the TMA loads of z (default 1)
 *   "cl_tma"               channels-last kernels: tensor-map TMA pipelines (1, default) or per-thread loads (0)
 *   "cl_tma_launches"      read-only counter of channels-last calls that took the TMA kernels (setting it resets it)
 *   "tail_stamps"          1: the in-kernel tail writes 16 phase clocks into the last 128 bytes of the workspace
 *   "wavelet_resident"     0 makes wtpse_wavelet_resident_cluster report 0 for every shape (per-level kernels)
 *   "wavelet_split"        fused-plan choice: -1 automatic, 0 whole map resident whenever it fits, 1 level 1 streamed
 *   "wavelet_tiles"        level 1 of the streamed plan as TMA pipelines (1, default) or per-thread loads (0)
 *   "wavelet_db2"          db2 streamed levels: factored one-/two-level passes of wavelet_db2.cu (1, default) or the round-1 level kernels (0)
 *   "wavelet_db2_two"      0: the factored passes take one level each (LL1 goes through memory)
 *   "wavelet_db2_deep"     1: keep peeling levels with the pass kernels as long as the band's width allows (default 0: see the planner)
 *   "wavelet_db2_rf" / "_ri" / "_nw2"   overrides of the pass geometry (strip rows, piece rows, warps of the level-2 group; 0 = automatic)
 *   "wavelet_haar_passes"  Haar levels through the pass kernels (1, default) or the band kernel / round-1 level kernels only (0)
 *   "wavelet_haar_min_log2px"  smallest map (log2 of its pixels, default 16) for which Haar takes the pass kernels
 *   "wavelet_peel_max"     most levels streamed before the resident stage (default 8)
 *   "wavelet_cluster_max"  largest cluster size of the resident stage (1..8, default 8)
 * Return WTPSE_OK, or WTPSE_ERR_INVALID for an unknown name. */
int wtpse_debug_set(const char* name, int value);
int wtpse_debug_get(const char* name, int* value);

/* ---- stress helper for the programmatic-dependent-launch paths ---------------------------------
 * Enqueues a kernel that signals griddepcontrol.launch_dependents at once, spins for spin_cycles clocks and only then copies
 * src -> dst (n floats).  A library kernel launched right behind it that read dst before its own griddepcontrol.wait would see
 * dst's old bytes (tests/test_gpu_fusion.py fills them with NaN). */
int wtpse_debug_pdl_slow_copy(float* dst, const float* src, long long n, long long spin_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WTPSE_B200_DEBUG_H */
