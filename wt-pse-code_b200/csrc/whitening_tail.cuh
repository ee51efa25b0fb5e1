// In-kernel forward tail of the Gram kernels: everything between the per-CTA partial Grams and the scalars
// (algorithms.py:1283-1307 after the bmm, compute_MMD.forward algorithms.py:102-121), executed by whichever CTA
// arrives LAST -- no second or third launch, no grid-wide wait:
//
//   per sample   every CTA that holds a partial Gram of sample b stores it, fences, and takes a ticket
//                (atomicAdd on ticket[b]).  The CTA that draws the last ticket sums the sample's partials in slot
//                order (fixed order: bit-reproducible, the atomics only count), forms f_cor = G/(P-1) + eps*I,
//                off_b, diag_b and the 120-d upper-triangle vector, and takes a ticket on ticket[B].
//   whole batch  the CTA that completes the last sample runs the all-to-all part in its (by then idle) pipeline
//                buffers: L_off, L_diag, pairwise expm1(-D), per-domain block sums, L_dom -- and, because the
//                backward needs nothing else from the MMD, the gradient of L_dom with respect to every vector entry
//                (`domgrad`), so the backward pass is ONE launch that combines it with the three upstream scalars.
//
// Nobody waits for anybody: a CTA that is not last simply continues (or exits).  The tickets are reset by their
// last taker, so a workspace whose ticket area was zero before the launch is zero again after it
// (include/wtpse_b200.h: wtpse_whitening_ticket_bytes).
//
// The arithmetic is the same code the stand-alone epilogue kernels run (mmd_device.cuh), in the same order: the two
// paths agree bit for bit (tests/test_gpu_parity.py).
#pragma once
#include "mmd_device.cuh"

namespace wtpse {

struct TailParams {
    int* ticket;             // [B + 1]: per-sample arrival counters, then the finished-sample counter
    const float* partial;    // [B][nslots][136]
    int nslots;
    int B;
    long long P;
    int n, K;
    float margin, eps;
    float* gram;             // [B][16][16]
    float* rowstat;          // [B][2]
    float* vd;               // [B][124] workspace: the upper-triangle vectors
    float* losses;           // [4]
    float* domgrad;          // [B][120]: d L_dom / d v_b (unscaled by the upstream gradient)
    long long* stamps;       // diagnostics (tools/tail_phases.py): 16 clock64() values of the CTA that finishes the batch, or nullptr
};

// phase timestamps: thread 0 of every CTA keeps them in registers; only the CTA that runs the whole-batch phase -- the
// critical path of the kernel's tail -- writes them out
struct TailClock {
    long long t[6];
    __device__ __forceinline__ void mark(int i) { t[i] = clock64(); }
};

// shared memory of the whole-batch phase: v [M][124] | U [M][M] | stat [B][2] | blk [K*K] f64 | coef [M][M] | wdom [K*K] | domk [M]
__host__ __device__ inline size_t tail_smem_bytes(int B, int M, int K) {
    const size_t kk = size_t(K > 0 ? K : 1) * size_t(K > 0 ? K : 1);
    return epi_mem_bytes(B, M, K) + (round4(size_t(M) * M) + round4(kk) + round4(size_t(M))) * sizeof(float);
}

// All NT threads of the calling group (named barrier `bar`) have stored what the ticket publishes.  Returns true, in
// every thread, iff this CTA drew ticket number expected - 1.
// Release/acquire through ONE thread (the pattern of CUTLASS' Semaphore): the barrier orders every thread's stores before
// thread 0's acq_rel atomic at GPU scope, which publishes them (release is cumulative over what the barrier
// synchronised) and, when it draws the last ticket, acquires the other CTAs' stores; the second barrier extends that to
// the whole group.  No 224-thread membar.gl on either side.
template <int NT>
__device__ __forceinline__ bool tail_take_ticket(int* counter, int expected, int tid, int* flag, int bar) {
    named_bar_sync(bar, NT);
    if (tid == 0) {
        int old;
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
        *flag = (old == expected - 1) ? 1 : 0;
    }
    named_bar_sync(bar, NT);
    return *flag != 0;
}

// Last-ticket holder of sample b: partial slots -> gram, rowstat, vd.  `cnt` slots are summed in slot order.
template <int NT>
__device__ __forceinline__ void tail_reduce_sample(const TailParams& tp, int b, int cnt, const IndexTables& tab, float* wred,
                                                   int tid, int bar) {
    const int warp = tid >> 5, lane = tid & 31;
    float off = 0.f, dg = 0.f;
    if (tid < kTri) {
        // same summation tree as gram_reduce_kernel (whitening_epilogue.cu): the allocated slots are cut into four
        // consecutive parts, each summed in slot order, then ((p0 + p1) + p2) + p3 -- identical bits on both paths
        const float* src = tp.partial + ((long long)b * tp.nslots) * kTri + tid;
        const int per = (tp.nslots + 3) / 4;
        float part[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k0 = q * per, k1 = (k0 + per < tp.nslots) ? k0 + per : tp.nslots;
            float acc = 0.f;
            for (int k = k0; k < k1; k += 8) {     // independent loads in flight, adds in slot order
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (k + u < k1 && k + u < cnt) ? __ldcg(src + (long long)(k + u) * kTri) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += v[u];
            }
            part[q] = acc;
        }
        const float s = ((part[0] + part[1]) + part[2]) + part[3];
        const int ij = tab.tri[tid], i = ij >> 4, j = ij & 15;
        float g = s / float(tp.P - 1);                           // .div(HW - 1), algorithms.py:1283
        if (i == j) {
            g += tp.eps;                                         // + eps * eye
            dg = fabsf(g - 1.0f);                                // |f_cor_masked_diag - I|, :1297
            tp.gram[b * 256 + i * kC + i] = g;
        } else {
            off = fabsf(g);                                      // |f_cor_masked|, :1289
            tp.gram[b * 256 + i * kC + j] = g;
            tp.gram[b * 256 + j * kC + i] = g;
            tp.vd[size_t(b) * kVStride + off_idx(i, j)] = g;
        }
    }
    if (warp < 5) {                                              // the 136 entry threads live in warps 0..4
        off = warp_sum(off);
        dg = warp_sum(dg);
        if (lane == 0) { wred[warp] = off; wred[8 + warp] = dg; }
    }
    named_bar_sync(bar, NT);
    if (tid == 0) {
        float so = 0.f, sd = 0.f;
        for (int w = 0; w < 5; ++w) { so += wred[w]; sd += wred[8 + w]; }
        tp.rowstat[b * 2 + 0] = so - tp.margin;
        tp.rowstat[b * 2 + 1] = sd - tp.margin;
        tp.ticket[b] = 0;                                        // leave the ticket area as we found it
    }
}

// Whole-batch phase, run by the CTA that completed the last sample.  `smem` holds tail_smem_bytes(B, M, K).
template <int NT>
__device__ __forceinline__ void tail_final(const TailParams& tp, float* smem, const IndexTables& tab, int tid, int bar,
                                           const TailClock& clk) {
    const int warp = tid >> 5, lane = tid & 31;
    const bool stamp = tp.stamps != nullptr && tid == 0;
    if (stamp) {
#pragma unroll
        for (int i = 0; i < 6; ++i) tp.stamps[i] = clk.t[i];
        tp.stamps[6] = clock64();
    }
    constexpr int kWarps = NT / 32;
    const int B = tp.B;
    const DomainInfo dom = make_domain(B, tp.n, tp.K);
    const int M = dom.M;
    const EpiMem mem = resolve_mem<true>(smem, nullptr, B, M);
    float* coef = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem) + epi_mem_bytes(B, M, tp.K));
    float* wdom = coef + round4(size_t(M) * M);        // [K][K]: the domain-pair weight of mmd_coefficient, computed once
    int* domk = reinterpret_cast<int*>(wdom + round4(size_t(tp.K > 0 ? tp.K : 1) * size_t(tp.K > 0 ? tp.K : 1)));   // [M]: domain of a sample (an integer division each, once)

    // A. stage the vectors and the row statistics the sample reducers left in global memory.  Every load of the first
    //    pass (8 vector pieces + one statistic per thread) is issued before anything is stored: one L2 round trip for
    //    M <= 48, B <= NT / 2.
    {
        const float4* src4 = reinterpret_cast<const float4*>(tp.vd);
        float4* dst4 = reinterpret_cast<float4*>(mem.v);
        const int n4 = M * (kVStride / 4);
        const float st0 = tid < 2 * B ? __ldcg(tp.rowstat + tid) : 0.f;
        for (int base = 0; base < n4; base += 8 * NT) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * NT + tid;
                v[u] = idx < n4 ? __ldcg(src4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * NT + tid;
                if (idx < n4) dst4[idx] = v[u];
            }
        }
        if (tid < 2 * B) mem.stat[tid] = st0;
        for (int idx = NT + tid; idx < 2 * B; idx += NT) mem.stat[idx] = __ldcg(tp.rowstat + idx);
        // mmd_coefficient(dom, a, c, E) = E * w(domain of a, domain of c) / npairs: tabulate w * (1 / 1) per domain pair
        // with the very expressions mmd_coefficient uses, so the product below has its bits
        for (int a = tid; a < M; a += NT) domk[a] = dom.domain_of(a);
        for (int idx = tid; idx < tp.K * tp.K; idx += NT) {
            const int ka = idx / tp.K, kc = idx - ka * tp.K;
            float w;
            if (ka == kc) {
                const float nk = float(dom.size(ka));
                w = -2.0f * float(dom.K - 1) / (nk * nk);
            } else {
                w = 2.0f / (float(dom.size(ka)) * float(dom.size(kc)));
            }
            wdom[idx] = w;
        }
    }
    named_bar_sync(bar, NT);
    if (stamp) tp.stamps[7] = clock64();

    // B. pairwise u = exp(-D) - 1 (the loss) and the gradient coefficients (the backward's seed)
    {
        float* U = mem.U;
        const float npairs = float(dom.K) * float(dom.K - 1) * 0.5f;
        const int K = tp.K;
        pairwise_upper_n(mem.v, M, tid, NT, [U, coef, wdom, domk, M, K, npairs](int a, int c, float D) {
            const float u = expm1f(-D);
            U[a * M + c] = u;
            U[c * M + a] = u;
            // == mmd_coefficient(dom, a, c, expf(-D)), which is symmetric in (a, c)
            const float cf = expf(-D) * wdom[domk[a] * K + domk[c]] / npairs;
            coef[a * M + c] = cf;
            coef[c * M + a] = cf;
        });
        for (int a = tid; a < M; a += NT) {
            U[a * M + a] = expm1f(-1e-30f);                      // D(a,a) = 0 -> clamp_min_(1e-30)
            coef[a * M + a] = 0.f;
        }
    }
    named_bar_sync(bar, NT);
    if (stamp) tp.stamps[8] = clock64();

    // C. per-domain-pair block sums (all warps, one block each for K = 3)
    if (M > 0) domain_block_sums(mem.U, dom, mem.blk, 0, kWarps, warp, lane);
    named_bar_sync(bar, NT);
    if (stamp) tp.stamps[9] = clock64();
    // D. Two independent strands:
    //   warp 0        the scalars: instance terms and L_dom (float64 combine: a latency chain of one warp);
    //   warps 1 ..    d L_dom / d v_b[o] for every MMD sample (`domgrad`), register-blocked over 5 samples per thread so that
    //                 every piece of v is read from shared memory once per block instead of once per sample
    //                 (M = 30: one round of tasks for the six or five warps there are).
    // (Before: every warp in both, one after the other, one thread per (sample, piece): 4.1 + 1.1 us, the first bound by
    // shared-memory bandwidth -- 0.5 MB through one SM.)
    if (warp == 0) {
        float so = 0.f, sd = 0.f;
        for (int b = lane; b < B; b += 32) {
            so += clamp0(mem.stat[b * 2 + 0] / float(kOff));     // clamp(off_diag_sum / 120, min=0), :1290
            sd += clamp0(mem.stat[b * 2 + 1] / float(kC));       // clamp(diag_sum / 16, min=0), :1298
        }
        so = warp_sum(so) / float(B);
        sd = warp_sum(sd) / float(B);
        const float pen = mmd_from_blocks(mem.blk, dom, lane);
        if (lane == 0) {
            tp.losses[0] = so;
            tp.losses[1] = sd;
            tp.losses[3] = so + sd;
            tp.losses[2] = pen;
            tp.ticket[B] = 0;
        }
    } else {
        constexpr int kBlk = NT >= 224 ? 5 : 6;       // M = 30: 6 x 30 = 180 tasks on 192 threads, or 5 x 30 = 150 on 160
        const int nblk = (M + kBlk - 1) / kBlk;
        for (int t = tid - 32; t < nblk * (kOff / 4); t += NT - 32) {
            const int blk = t / (kOff / 4), q = t - blk * (kOff / 4);
            mmd_grad_block4<kBlk>(mem.v, coef, M, blk * kBlk, q, tp.domgrad);
        }
        if (tp.stamps != nullptr && tid == 32) tp.stamps[10] = clock64();
    }
}

// Convenience: called by the NT worker threads right after they stored sample b's partial into its slot.
// `expected` = number of partials sample b receives in this launch.  `smem_final` may be the pipeline buffers.
// clk.t[0] / t[1]: set by the caller before / after it flushed its partial.
template <int NT>
__device__ __forceinline__ void tail_after_flush(const TailParams& tp, int b, int expected, const IndexTables& tab, float* wred,
                                                 int* flag, float* smem_final, int tid, int bar, TailClock& clk) {
    const bool last_of_sample = tail_take_ticket<NT>(tp.ticket + b, expected, tid, flag, bar);
    clk.mark(2);
    if (!last_of_sample) return;
    tail_reduce_sample<NT>(tp, b, expected, tab, wred, tid, bar);
    clk.mark(3);
    const bool last_of_batch = tail_take_ticket<NT>(tp.ticket + tp.B, tp.B, tid, flag, bar);
    clk.mark(4);
    if (!last_of_batch) return;
    tail_final<NT>(tp, smem_final, tab, tid, bar, clk);
}

}  // namespace wtpse
