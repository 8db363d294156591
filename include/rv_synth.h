/*
 * rv_synth.h — the seeded synthetic hand workload of BASELINE.json configs[1] (SURVEY.md §8 d "Config 2").
 *
 * Input data, not an algorithm under test: one definition, compiled for the device (rv_hand_queries_seeded fills a device
 * buffer, so the 10^7-hand benchmark needs no 560 MB upload) and for the host (the oracle and the tests build the same
 * queries on the CPU).  Hand h of the stream:
 *   x0 = splitmix64(0xA6A21 + h), x_{i+1} = splitmix64(x_i) supply the random numbers;
 *   h % 10 != 9  uniform stratum: a Fisher-Yates prefix of the 136 tile ids draws 14 distinct tiles; the win tile is the
 *                last one drawn; the next tile drawn is the dora indicator;
 *   h % 10 == 9  positive stratum: four random mentsu (sequence or triplet) and a pair, at most four copies per kind
 *                (re-drawn up to 8 times, else the hand falls back to the uniform stratum), copies taken in id order, the win
 *                tile one of the 14 at random, the dora indicator a random tile id;
 *   context      c = bits 20-24 of splitmix64(h) (h % 32 would tie the context to the stratum: h % 10 == 9 is always odd):
 *                tsumo = c & 1, riichi = (c >> 1) & 1, seat wind = (c >> 2) & 3, round wind = (c >> 4) & 1.
 * Uniformly random hands are almost never complete, so the positive stratum is what exercises yaku / fu / score.
 */
#ifndef RV_SYNTH_H
#define RV_SYNTH_H

#include <stdint.h>

#include "riichienv_b200.h"

#ifdef __CUDACC__
#define RV_SYNTH_FN static __host__ __device__ __forceinline__
#else
#define RV_SYNTH_FN static inline
#endif

RV_SYNTH_FN uint64_t rv_synth_mix(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

RV_SYNTH_FN void rv_synth_hand(uint64_t h, rv_hand_query* q) {
  uint64_t x = rv_synth_mix(0xA6A21ull + h);
  uint8_t* raw = (uint8_t*)q;
  for (unsigned i = 0; i < sizeof(rv_hand_query); i++) raw[i] = 0;
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 4; k++) q->meld_tiles[i][k] = RV_NONE;
  for (int i = 0; i < 5; i++) q->dora_ind[i] = q->ura_ind[i] = RV_NONE;
  const unsigned c = (unsigned)(rv_synth_mix(h) >> 20) & 31;
  q->cond = (uint16_t)((c & 1 ? RV_C_TSUMO : 0) | (c & 2 ? RV_C_RIICHI : 0));
  q->player_wind = (uint8_t)((c >> 2) & 3);
  q->round_wind = (uint8_t)((c >> 4) & 1);
  q->n_tiles = 14;
  q->n_dora = 1;
  int done = 0;
  if (h % 10 == 9) {
    for (int attempt = 0; attempt < 8 && !done; attempt++) {
      uint8_t cnt[34];
      for (int k = 0; k < 34; k++) cnt[k] = 0;
      int ok = 1;
      for (int m = 0; m < 4; m++) {
        x = rv_synth_mix(x);
        if (x & 1) {
          int kind = (int)((x >> 8) % 34);
          cnt[kind] += 3;
        } else {
          int suit = (int)((x >> 8) % 3), start = (int)((x >> 16) % 7);
          for (int d = 0; d < 3; d++) cnt[9 * suit + start + d] += 1;
        }
      }
      x = rv_synth_mix(x);
      cnt[(int)((x >> 8) % 34)] += 2;
      for (int k = 0; k < 34; k++)
        if (cnt[k] > 4) ok = 0;
      if (!ok) continue;
      int n = 0;
      for (int k = 0; k < 34; k++)
        for (int j = 0; j < cnt[k]; j++) q->tiles[n++] = (uint8_t)(4 * k + j);
      x = rv_synth_mix(x);
      q->win_tile = q->tiles[(int)((x >> 8) % 14)];
      x = rv_synth_mix(x);
      q->dora_ind[0] = (uint8_t)((x >> 8) % 136);
      done = 1;
    }
  }
  if (!done) {
    uint8_t deck[136];
    for (int i = 0; i < 136; i++) deck[i] = (uint8_t)i;
    for (int i = 0; i < 15; i++) {
      x = rv_synth_mix(x);
      int j = i + (int)((x >> 8) % (uint64_t)(136 - i));
      uint8_t t = deck[i];
      deck[i] = deck[j];
      deck[j] = t;
    }
    for (int i = 0; i < 14; i++) q->tiles[i] = deck[i];
    q->win_tile = deck[13];
    q->dora_ind[0] = deck[14];
  }
}

#endif
