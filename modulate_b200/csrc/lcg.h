// lcg.h -- modular arithmetic of the keystream recurrence, shared by host and device code.
//
// The reference steps its key with Schrage's method (CEncryptionCycler.cpp:16-25), which is
// x' = 16807 * x mod (2^31 - 1) with residue 0 held as m.  Everything here is that recurrence
// re-expressed for a machine with a 32x32->64 multiplier: a Mersenne fold instead of divisions,
// and an O(log n) jump-ahead instead of n sequential steps.  No code is shared with the reference.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define MODLCG_HD __host__ __device__ __forceinline__ constexpr
#else
#define MODLCG_HD inline constexpr
#endif

namespace modlcg {

constexpr uint32_t kM = 0x7FFFFFFFu;  // 2^31 - 1 (prime)
constexpr uint32_t kA = 16807u;       // primitive root of kM
constexpr uint32_t kA2 = 2u * kA;     // the step multiplies by 2a so the product splits at bit 31

// x * y mod m for x, y in [0, m].  Result in [0, m]; it is 0 only when x*y == 0 and m (== 0 mod m)
// only when an operand was m, so operands in [0, m-1] give a canonical result in [0, m-1].
//   x * 2y = hi * 2^32 + lo  =>  x*y = hi * 2^31 + (lo >> 1)  ==  hi + (lo >> 1)   (mod m)
// hi + (lo >> 1) <= 2^32 - 2, so one more end-around fold lands in [0, m].
MODLCG_HD uint32_t mulmod(uint32_t x, uint32_t y)
{
    const uint64_t p = (uint64_t)x * (uint64_t)(y << 1);
    const uint32_t t = (uint32_t)(p >> 32) + ((uint32_t)p >> 1);
    return (t & kM) + (t >> 31);
}

// One keystream step on a LAZILY reduced state: s < 2^31 + 2^15 (not necessarily < m).
// Result obeys the same bound (hi <= 16807), so the chain never needs a conditional subtract.
MODLCG_HD uint32_t step_lazy(uint32_t s)
{
    const uint64_t p = (uint64_t)s * (uint64_t)kA2;
    return (uint32_t)(p >> 32) + ((uint32_t)p >> 1);
}

// Low 8 bits of the canonical residue of a lazily reduced state that is 0 or != 0 (mod m).
// m == -1 (mod 256), so subtracting m when bit 31 is set is adding 1 in the low byte.  (A lazily
// reduced nonzero residue is never exactly m, so "bit 31 set" is "s > m".)
MODLCG_HD uint32_t low8_canonical(uint32_t s) { return s + (s >> 31); }

// Signed 32-bit key -> NEGATED residue n0 = -k0 mod m in [0, m-1].
//
// The kernels run the recurrence on n_i = -state_i: it obeys the same x' = a*x law, and
//     low8(canonical(state_i)) ^ 0xFF = 255 - low8(state_i) = low8(m - state_i) = low8(canonical(n_i))
// (m == 255 mod 256), so the reference's "^ 0xFF" costs nothing.  The identity stream (k0 == 0,
// which the reference holds at state == m) is n == 0, a fixed point of every multiply here.
// INT_MIN, -1 and m-1 all have k0 = m-1, i.e. n0 = 1.
MODLCG_HD uint32_t key_to_neg_state(int32_t key)
{
    int32_t r = key % (int32_t)kM;  // C remainder: sign follows the dividend
    if (r < 0)
        r += (int32_t)kM;           // r = k0 in [0, m-1]
    return r == 0 ? 0u : kM - (uint32_t)r;
}

// Mathematical residue k0 of a signed key, in [0, m-1].
MODLCG_HD uint32_t key_residue(int32_t key)
{
    int32_t r = key % (int32_t)kM;
    if (r < 0)
        r += (int32_t)kM;
    return (uint32_t)r;
}

// a^e mod m by square-and-multiply; the exponent lives mod (m-1) since a is a primitive root.
MODLCG_HD uint32_t pow_a(uint64_t e)
{
    e %= (uint64_t)(kM - 1u);
    uint32_t result = 1u, base = kA;
    while (e) {
        if (e & 1u)
            result = mulmod(result, base);
        base = mulmod(base, base);
        e >>= 1;
    }
    return result;
}

// a^(-e) mod m.
MODLCG_HD uint32_t pow_a_inv(uint64_t e)
{
    e %= (uint64_t)(kM - 1u);
    return pow_a((uint64_t)(kM - 1u) - e);
}

}  // namespace modlcg
