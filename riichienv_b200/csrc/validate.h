// Host-side sanity check of a record handed in through rv_vec_set_state (plain C++, no CUDA).
#pragma once
#include <stdint.h>

#include "../../include/riichienv_b200.h"

// A record handed in through the API must keep every counter inside its array and every tile id inside 0..135: the device
// code indexes fixed arrays with them.  More than four copies of a kind in a hand are tolerated (table indices are clamped,
// hand.cuh suit_key) up to the 15 a 4-bit histogram cell can count.
inline const char* rv_state_defect(const rv_game_state& s) {
  const int np = s.game_mode >= 3 ? 3 : 4;
  if (s.game_mode > 5) return "game_mode must be 0..5";
  if (s.wall_len != 0 && s.wall_len != (np == 3 ? 108 : 136)) return "wall_len does not match the game mode";   // 0: never dealt
  if (s.wall_top > 136 || s.rinshan_draw_count > s.wall_top) return "wall cursors out of range";
  if (s.n_dora > 5) return "more than 5 dora indicators";
  if (s.phase > 1) return "phase must be 0 or 1";
  if (s.current_player >= np && s.current_player != RV_NONE) return "current_player out of range";   // RV_NONE: nobody to act (event-driven records)
  if (s.oya >= np) return "oya out of range";
  for (int p = 0; p < RV_NP; p++) {
    if (s.hand_len[p] > RV_HAND_CAP) return "a hand holds at most 16 tiles";
    if (s.n_melds[p] > 4) return "a seat holds at most 4 melds";
    if (s.n_river[p] > RV_RIVER_CAP) return "a river holds at most 32 discards";
    if (s.n_claims[p] > RV_MAX_CLAIMS) return "too many claims";
    uint8_t cnt[34] = {0};
    for (int k = 0; k < s.hand_len[p]; k++) {
      if (s.hand[p][k] >= 136) return "hand tile id out of range";
      if (++cnt[s.hand[p][k] >> 2] > 15) return "a hand holds more than 15 tiles of one kind";
    }
    for (int m = 0; m < s.n_melds[p]; m++) {
      if (s.meld_type[p][m] > RV_MELD_KAKAN) return "meld type out of range";
      for (int k = 0; k < 3; k++)
        if (s.meld_tiles[p][m][k] >= 136) return "meld tile id out of range";
      if (s.meld_tiles[p][m][3] >= 136 && s.meld_tiles[p][m][3] != RV_NONE) return "meld tile id out of range";
    }
    for (int k = 0; k < s.n_river[p]; k++)
      if (s.river[p][k] >= 136) return "river tile id out of range";
  }
  if (s.drawn_tile >= 136 && s.drawn_tile != RV_NONE) return "drawn_tile out of range";
  if (s.last_discard_tile >= 136 && s.last_discard_tile != RV_NONE) return "last_discard tile out of range";
  if (s.last_discard_pid >= np && s.last_discard_pid != RV_NONE) return "last_discard seat out of range";
  for (int i = s.rinshan_draw_count; i < s.wall_top; i++)
    if (s.wall[i] >= 136) return "wall tile id out of range";
  return nullptr;
}
