// Track W: per-site arithmetic of one analysis / synthesis level, shared by the cluster-resident kernel
// (wavelet_resident.cu, operands in shared memory) and the streaming level-1 kernels (wavelet_stream.cu, operands in
// global memory).  A thread always owns TWO horizontally adjacent half-resolution sites and walks down the rows, so
// the row-filtered values (analysis) / column-synthesised values (synthesis) of the TAPS-2 overlapping rows are
// computed once and carried in registers.  PARITY UNPINNED (oracle/wavelet_np.py is this repository's own spec).
#pragma once
#include "wavelet_bank.cuh"

namespace wtpse {

// sign(d) of a detail coefficient as a 2-bit code (0: negative, 1: zero, 2: positive); the three sub-bands of a site
// share one byte: LH | HL << 2 | HH << 4.  The L1 loss needs nothing else from the detail bands for its gradient.
__device__ __forceinline__ unsigned sign_code(float v) { return v > 0.f ? 2u : (v < 0.f ? 0u : 1u); }
// 2-bit code -> -1 / 0 / +1 without an integer-to-float conversion: 0x4B400000 is 1.5 * 2^23
__device__ __forceinline__ float sign_value(unsigned b, int shift) {
    return __uint_as_float(0x4B400000u + ((b >> shift) & 3u)) - 12582913.0f;
}
__device__ __forceinline__ unsigned sign_pack2(const float2& LH, const float2& HL, const float2& HH) {
    const unsigned ca = sign_code(LH.x) | (sign_code(HL.x) << 2) | (sign_code(HH.x) << 4);
    const unsigned cb = sign_code(LH.y) | (sign_code(HL.y) << 2) | (sign_code(HH.y) << 4);
    return ca | (cb << 8);
}

// low/high-pass along W of one input row: x[0..TAPS+1] are the input columns 4jj .. 4jj+TAPS+1 (periodic).
// Scalar FFMA on purpose: the packed fma.rn.f32x2 form (two sites per instruction, 11 % fewer instructions in the
// level-1 analysis kernel) was measured SLOWER on B200 -- FFMA2 holds the FMA pipe for two cycles, issue utilisation
// fell from 65 % to 54 % and the kernel went from 31.1k to 33.7k active cycles.
template <int TAPS>
__device__ __forceinline__ void row_filter(const float (&x)[TAPS + 2], float& lo0, float& hi0, float& lo1, float& hi1) {
    float s0 = 0.f, d0 = 0.f, s1 = 0.f, d1 = 0.f;
#pragma unroll
    for (int l = 0; l < TAPS; ++l) {
        s0 = fmaf(Bank<TAPS>::h(l), x[l], s0); d0 = fmaf(Bank<TAPS>::g(l), x[l], d0);
        s1 = fmaf(Bank<TAPS>::h(l), x[l + 2], s1); d1 = fmaf(Bank<TAPS>::g(l), x[l + 2], d1);
    }
    lo0 = s0; hi0 = d0; lo1 = s1; hi1 = d1;
}

// low/high-pass along H over the TAPS carried rows -> the four sub-band values of both sites
template <int TAPS>
__device__ __forceinline__ void col_filter(const float (&lo0)[TAPS], const float (&hi0)[TAPS], const float (&lo1)[TAPS],
                                           const float (&hi1)[TAPS], float2& LL, float2& LH, float2& HL, float2& HH) {
    LL = make_float2(0.f, 0.f); LH = LL; HL = LL; HH = LL;
#pragma unroll
    for (int k = 0; k < TAPS; ++k) {
        LL.x = fmaf(Bank<TAPS>::h(k), lo0[k], LL.x); LL.y = fmaf(Bank<TAPS>::h(k), lo1[k], LL.y);
        LH.x = fmaf(Bank<TAPS>::h(k), hi0[k], LH.x); LH.y = fmaf(Bank<TAPS>::h(k), hi1[k], LH.y);
        HL.x = fmaf(Bank<TAPS>::g(k), lo0[k], HL.x); HL.y = fmaf(Bank<TAPS>::g(k), lo1[k], HL.y);
        HH.x = fmaf(Bank<TAPS>::g(k), hi0[k], HH.x); HH.y = fmaf(Bank<TAPS>::g(k), hi1[k], HH.y);
    }
}

__device__ __forceinline__ float abs_sum(const float2& LH, const float2& HL, const float2& HH) {
    return fabsf(LH.x) + fabsf(LH.y) + fabsf(HL.x) + fabsf(HL.y) + fabsf(HH.x) + fabsf(HH.y);
}

// Column synthesis of one coefficient row for the output sites A = 2q and B = 2q+1.  l01 / b01: low-low gradient and
// packed signs of columns 2q, 2q+1; lm / bm: of column 2q-1 (periodic; only db2 reads it).  The detail gradient of a
// site is sc * sign.  tL / tH [A parity 0, A parity 1, B parity 0, B parity 1]: low-row / high-row content of the
// four output columns 4q .. 4q+3.  Every output is one multiply + a chain of FMAs (4 terms for db2, 2 for Haar).
template <int TAPS, bool kHasLL = true>
__device__ __forceinline__ void col_synth_vals(const float2& l01, float lm, unsigned b01, unsigned bm, float sc, float (&tL)[4],
                                               float (&tH)[4]) {
    const float lh0 = sign_value(b01, 0), hl0 = sign_value(b01, 2), hh0 = sign_value(b01, 4);
    const float lh1 = sign_value(b01, 8), hl1 = sign_value(b01, 10), hh1 = sign_value(b01, 12);
#pragma unroll
    for (int pc = 0; pc < 2; ++pc) {
        const float h = Bank<TAPS>::h(pc), hs = Bank<TAPS>::h(pc) * sc, gs = Bank<TAPS>::g(pc) * sc;
        tL[pc] = kHasLL ? fmaf(h, l01.x, gs * lh0) : gs * lh0;
        tH[pc] = fmaf(hs, hl0, gs * hh0);
        tL[2 + pc] = kHasLL ? fmaf(h, l01.y, gs * lh1) : gs * lh1;
        tH[2 + pc] = fmaf(hs, hl1, gs * hh1);
    }
    if (TAPS == 4) {
        const float lhm = sign_value(bm, 0), hlm = sign_value(bm, 2), hhm = sign_value(bm, 4);
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
            const float h = Bank<TAPS>::h(2 + pc), hs = Bank<TAPS>::h(2 + pc) * sc, gs = Bank<TAPS>::g(2 + pc) * sc;
            // second tap: site A (column 2q) takes it from column 2q-1, site B (2q+1) from column 2q.  Two short
            // independent chains per output instead of one 4-deep chain: the single chain was measured slower
            // (issue utilisation 74 % -> 63 % in the level-1 synthesis kernel).
            tL[pc] += kHasLL ? fmaf(h, lm, gs * lhm) : gs * lhm;
            tH[pc] += fmaf(hs, hlm, gs * hhm);
            tL[2 + pc] += kHasLL ? fmaf(h, l01.x, gs * lh0) : gs * lh0;
            tH[2 + pc] += fmaf(hs, hl0, gs * hh0);
        }
    }
}

// Row synthesis: output row 2i + pr, times gs, from the column-synthesised coefficient rows i (c) and i-1 (p)
template <int TAPS>
__device__ __forceinline__ void row_synth(const float (&cL)[4], const float (&cH)[4], const float (&pL)[4], const float (&pH)[4],
                                          int pr, float gs, float (&o)[4]) {
    const float h0 = Bank<TAPS>::h(pr) * gs, g0 = Bank<TAPS>::g(pr) * gs;
    const float h1 = TAPS == 4 ? Bank<TAPS>::h(TAPS == 4 ? 2 + pr : 0) * gs : 0.f;
    const float g1 = TAPS == 4 ? Bank<TAPS>::g(TAPS == 4 ? 2 + pr : 0) * gs : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        o[k] = fmaf(h0, cL[k], g0 * cH[k]);
        if (TAPS == 4) o[k] += fmaf(h1, pL[k], g1 * pH[k]);
    }
}

}  // namespace wtpse
