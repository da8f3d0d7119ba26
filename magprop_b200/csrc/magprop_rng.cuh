// magprop_rng.cuh -- counter-based random numbers of the ensemble move: Philox4x32-10 draws and the keyed
// permutation that defines the per-step random halves.  Plain C++ apart from the qualifiers, so the CPU
// tests can compile it for the host and pin it against their NumPy restatement.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MP_RNG_HD __host__ __device__ inline
#else
#define MP_RNG_HD inline
#endif

namespace mp {

// ---- counter-based RNG: Philox4x32-10 (Salmon et al. 2011) -------------------
struct Philox {
  uint32_t c[4];
};
MP_RNG_HD Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox o;
  o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
  return o;
}
// 53-bit uniform in (0,1): never 0 so log() is finite
MP_RNG_HD double u01(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// ---- the ensemble order: a keyed pseudo-random permutation of [0, n) -----------------------------
// emcee's RedBlueMove re-draws the two halves every step (`inds = arange(n) % 2; random.shuffle(inds)`,
// SURVEY.md appendix C).  Here position g of the ensemble order holds walker P(g) and the halves are
// g < n/2 and g >= n/2; P is a balanced Feistel network on 2*hb bits (4^hb >= n) walked until it lands
// inside [0, n) -- a bijection any thread of any rank evaluates for one index in a few dozen integer
// instructions, so the split needs no index arrays, no sort and no communication.  Keys: one Philox block
// of (seed, step).  randomize = 0: P = identity (fixed halves).
struct SplitPerm {
  uint32_t n, k0, k1, mask;
  int hb, randomize;
};
MP_RNG_HD uint32_t mix32(uint32_t x) {   // an avalanching 32-bit hash ("lowbias32")
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
MP_RNG_HD uint32_t perm_at(const SplitPerm& p, uint32_t g) {
  if (!p.randomize) return g;
  uint32_t x = g;
  do {
    uint32_t L = x >> p.hb, R = x & p.mask;
    for (uint32_t r = 0; r < 6u; ++r) {
      const uint32_t F = mix32((R + r * 0x9E3779B9u) ^ ((r & 1u) ? p.k1 : p.k0)) & p.mask;
      const uint32_t nL = R;
      R = L ^ F;
      L = nL;
    }
    x = (L << p.hb) | R;
  } while (x >= p.n);
  return x;
}
// The inverse: at which position of the ensemble order walker w sits.  (Cycle-walking is symmetric: undo the
// rounds, and walk on while the result is outside [0, n).)
MP_RNG_HD uint32_t perm_inv(const SplitPerm& p, uint32_t w) {
  if (!p.randomize) return w;
  uint32_t x = w;
  do {
    uint32_t L = x >> p.hb, R = x & p.mask;
    for (int r = 5; r >= 0; --r) {
      const uint32_t F = mix32((L + (uint32_t)r * 0x9E3779B9u) ^ ((r & 1) ? p.k1 : p.k0)) & p.mask;
      const uint32_t pR = L;
      L = R ^ F;
      R = pR;
    }
    x = (L << p.hb) | R;
  } while (x >= p.n);
  return x;
}

inline SplitPerm make_split_perm(int n, uint64_t seed, uint64_t step, int randomize) {
  SplitPerm p;
  p.n = (uint32_t)n;
  p.randomize = randomize ? 1 : 0;
  p.hb = 1;
  while ((1ull << (2 * p.hb)) < (unsigned long long)n) ++p.hb;
  p.mask = (1u << p.hb) - 1u;
  const Philox k = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 0u, 2u, (uint32_t)seed, (uint32_t)(seed >> 32));
  p.k0 = k.c[0];
  p.k1 = k.c[1];
  return p;
}

}  // namespace mp
