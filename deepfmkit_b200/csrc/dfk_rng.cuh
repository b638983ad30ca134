// Counter-based random numbers for the device generators: Philox4x32-10 and fp32 Box-Muller normals.
// A draw is a pure function of (key, counter), so any part of a record can be produced anywhere.
#pragma once
#include "dfk_common.cuh"

namespace dfk {

DFK_D void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two standard normals from two 32-bit draws (Box-Muller).  The hardware approximations (lg2.approx, sin/cos.approx on
// an angle in [-pi, pi): absolute error ~2^-21) are six digits below the draw itself and leave no branch in the
// generator's inner loop; the radius is exact to fp32 rounding.
DFK_D void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
    const float ang = (static_cast<float>(b) * 2.3283064365386963e-10f - 0.5f) * 6.283185307179586f;  // [-pi, pi]
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cn;
    __sincosf(ang, &sn, &cn);
    z0 = rad * cn;
    z1 = rad * sn;
}

// Four standard normals from one Philox block (key, counter): two Box-Muller pairs in fp32.  The noise these feed
// is 20..60 dB below the signal, so 2^-24 granularity is seven digits below anything a fit can see; |z| <= 6.66.
DFK_D void normal4(unsigned long long key, unsigned long long counter, uint32_t stream, float z[4]) {
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), stream, 0u,
                  static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        box_muller(r[2 * h], r[2 * h + 1], z[2 * h], z[2 * h + 1]);
    }
}

}  // namespace dfk
