// ORACLE — TEST INFRASTRUCTURE ONLY (see hand.hpp header).
//
// Seeded wall shuffle.  Follows state/wall.rs:36-67 (splitmix64 -> StdRng::seed_from_u64
// -> slice.shuffle -> reverse).  The RNG itself lives in crates that are NOT vendored
// under /root/reference: rand 0.10.0, rand_core 0.10.0, chacha20 0.10.0
// (Cargo.lock:84-86,667-681).  Their published algorithm is restated here:
//   * SeedableRng::seed_from_u64: PCG32 (XSH-RR) stream expands the u64 into a 32-byte key
//   * StdRng = ChaCha, 12 rounds, 64-bit block counter (words 12,13) from 0, stream 0,
//     output consumed as consecutive little-endian u32 words
//   * SliceRandom::shuffle = partial_shuffle(len) with IncreasingUniform chunking
//   * random_range(..bound) for u32 = widening multiply with one bias-correction draw
// PARITY PARTLY PINNED at this boundary: no reference test pins (seed -> wall)
// (tests/env/test_riichienv.py:56-58 declines to), and the crates cannot be built here.
// What IS pinned (tests/test_oracle_golden.py::test_chacha_known_answers): the ChaCha block
// function and word order on the published ChaCha20/ChaCha12 zero-key keystreams and on the
// constant of rand's own `test_stdrng_construction` (StdRng::from_seed -> next_u64).  The
// PCG32 seed expansion and the chunked shuffle are restated from the crate sources, unpinned.
// Everything is isolated in wall_from_seed() so it can be corrected in one place.
#pragma once
#include <cstdint>
#include <vector>

namespace orc {

// state/wall.rs:83-88
inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct ChaCha12 {
  uint32_t key[8];
  uint64_t counter = 0;
  uint32_t buf[16];
  int idx = 16;
  int double_rounds = 6;  // StdRng: 12 rounds; the known-answer tests also run 20

  static inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
  static inline void qr(uint32_t* s, int a, int b, int c, int d) {
    s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 16);
    s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 12);
    s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 8);
    s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 7);
  }
  // rand_core SeedableRng::seed_from_u64 (PCG32 expansion)
  explicit ChaCha12(uint64_t state) {
    for (int i = 0; i < 8; i++) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
      uint32_t rot = (uint32_t)(state >> 59);
      key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
  }
  // raw 32-byte key (SeedableRng::from_seed) — used by the known-answer tests
  ChaCha12(const uint32_t* k, int rounds) : double_rounds(rounds / 2) {
    for (int i = 0; i < 8; i++) key[i] = k[i];
  }
  void refill() {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = (uint32_t)counter;
    in[13] = (uint32_t)(counter >> 32);
    in[14] = 0;
    in[15] = 0;
    uint32_t s[16];
    for (int i = 0; i < 16; i++) s[i] = in[i];
    for (int r = 0; r < double_rounds; r++) {
      qr(s, 0, 4, 8, 12); qr(s, 1, 5, 9, 13); qr(s, 2, 6, 10, 14); qr(s, 3, 7, 11, 15);
      qr(s, 0, 5, 10, 15); qr(s, 1, 6, 11, 12); qr(s, 2, 7, 8, 13); qr(s, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) buf[i] = s[i] + in[i];
    counter++;
    idx = 0;
  }
  uint32_t next_u32() {
    if (idx >= 16) refill();
    return buf[idx++];
  }
};

// rand::distr::uniform  UniformInt<u32>::sample_single_inclusive (0, bound-1)
inline uint32_t random_below(ChaCha12& rng, uint32_t range) {
  uint64_t m = (uint64_t)rng.next_u32() * range;
  uint32_t result = (uint32_t)(m >> 32), lo = (uint32_t)m;
  if (lo > (uint32_t)(0u - range)) {
    uint64_t m2 = (uint64_t)rng.next_u32() * range;
    uint32_t new_hi = (uint32_t)(m2 >> 32);
    if ((uint64_t)lo + new_hi > 0xFFFFFFFFull) result += 1;
  }
  return result;
}

// rand::seq::IncreasingUniform + SliceRandom::shuffle
template <class T>
inline void rand_shuffle(std::vector<T>& v, ChaCha12& rng) {
  uint32_t n = 0, chunk = 0;
  uint32_t chunk_remaining = 1;  // n == 0 -> first index is 0 without drawing
  for (size_t i = 0; i < v.size(); i++) {
    uint32_t next_n = n + 1;
    uint32_t rem;
    if (chunk_remaining == 0) {
      // calculate_bound_u32(next_n)
      uint32_t product = next_n, current = next_n + 1;
      while (true) {
        uint64_t p = (uint64_t)product * current;
        if (p > 0xFFFFFFFFull) break;
        product = (uint32_t)p;
        current++;
      }
      chunk = random_below(rng, product);
      rem = (current - next_n) - 1;
    } else {
      rem = chunk_remaining - 1;
    }
    uint32_t result;
    if (rem == 0) {
      result = chunk;
    } else {
      result = chunk % next_n;
      chunk /= next_n;
    }
    chunk_remaining = rem;
    n = next_n;
    std::swap(v[i], v[result]);
  }
}

// state/wall.rs:36-58 (4P) and state_3p/wall.rs:75-93 (3P tile set); returns the
// REVERSED order, i.e. the reference's `wall.tiles`.
inline std::vector<uint8_t> wall_from_seed(uint64_t seed, uint64_t hand_index, int n_tiles) {
  std::vector<uint8_t> w;
  if (n_tiles == 136) {
    for (int i = 0; i < 136; i++) w.push_back((uint8_t)i);
  } else {
    for (int i = 0; i < 136; i++) {
      int t34 = i / 4;
      if (t34 >= 1 && t34 <= 7) continue;
      w.push_back((uint8_t)i);
    }
  }
  ChaCha12 rng(splitmix64(seed + hand_index));
  rand_shuffle(w, rng);
  std::vector<uint8_t> r(w.rbegin(), w.rend());
  return r;
}

}  // namespace orc
