// Track W filter banks (orthonormal Haar / db2), shared by the per-level and the cluster-resident kernels.
// PARITY UNPINNED: conventions are this repository's own (oracle/wavelet_np.py).
#pragma once

namespace wtpse {

template <int TAPS>
struct Bank;
template <>
struct Bank<2> {
    __device__ static float h(int k) { return 0.70710678118654752f; }
    __device__ static float g(int k) { return k == 0 ? 0.70710678118654752f : -0.70710678118654752f; }
};
template <>
struct Bank<4> {
    // db2: h = [1+s3, 3+s3, 3-s3, 1-s3] / (4 sqrt2),  g[k] = (-1)^k h[3-k]
    __device__ static float h(int k) {
        return k == 0 ? 0.48296291314453414f : k == 1 ? 0.83651630373780790f : k == 2 ? 0.22414386804201339f : -0.12940952255126037f;
    }
    __device__ static float g(int k) {
        return k == 0 ? -0.12940952255126037f : k == 1 ? -0.22414386804201339f : k == 2 ? 0.83651630373780790f : -0.48296291314453414f;
    }
};

}  // namespace wtpse
