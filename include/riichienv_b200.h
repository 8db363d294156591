/*
 * riichienv_b200.h — C ABI of the B200-native batched Riichi mahjong simulator.
 *
 * This is the drop-in boundary for the reference's hot path (SURVEY.md §8 b).
 * The reference (smly/RiichiEnv) has no C ABI of its own: its boundary is the
 * PyO3 class `riichienv._riichienv.RiichiEnv` over `GameStateVariant`
 * (riichienv-python/src/env.rs:74-118).  Every entry point below names the
 * reference interface it replaces (file:line relative to /root/reference).
 * INTEGRATION.md shows the Rust `extern "C"` block + PyO3 shim a maintainer
 * would add on the reference side.
 *
 * Conventions: plain pointers and sizes, no torch / CUDA types.  Pointers named
 * `d_*` are DEVICE pointers (caller-allocated, e.g. torch tensors' data_ptr);
 * everything else is host memory.  All functions return 0 on success or a
 * negative RV_ERR_* code; rv_last_error() gives a message.  A handle is bound
 * to one CUDA device and one stream and is not thread-safe (one host thread +
 * one stream per GPU, as in the north star).
 */
#ifndef RIICHIENV_B200_H
#define RIICHIENV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------- */
#define RV_OK 0
#define RV_ERR_INVALID -1   /* bad argument (maps to PyValueError, env.rs:815-823) */
#define RV_ERR_CUDA -2      /* CUDA runtime failure / no device: the product has NO CPU fallback */
#define RV_ERR_UNSUPPORTED -3

/* ---- enums (values identical to the reference) ------------------------- */
/* action.rs:55-68 */
enum rv_action_type {
  RV_DISCARD = 0, RV_CHI = 1, RV_PON = 2, RV_DAIMINKAN = 3, RV_RON = 4, RV_RIICHI = 5,
  RV_TSUMO = 6, RV_PASS = 7, RV_ANKAN = 8, RV_KAKAN = 9, RV_KYUSHU_KYUHAI = 10, RV_KITA = 11,
  RV_NO_ACTION = 255 /* seat supplies no action this step (missing dict key in env.step) */
};
/* types.rs:57-63 */
enum rv_meld_type { RV_MELD_CHI = 0, RV_MELD_PON = 1, RV_MELD_DAIMINKAN = 2, RV_MELD_ANKAN = 3, RV_MELD_KAKAN = 4 };
/* action.rs:30-33 */
enum rv_phase { RV_WAIT_ACT = 0, RV_WAIT_RESPONSE = 1 };
/* env.rs:93-101 */
enum rv_game_mode {
  RV_4P_RED_SINGLE = 0, RV_4P_RED_EAST = 1, RV_4P_RED_HALF = 2,
  RV_3P_RED_SINGLE = 3, RV_3P_RED_EAST = 4, RV_3P_RED_HALF = 5
};
/* rule.rs:10-20 — one bit per GameRule field, in declaration order */
#define RV_RULE_RON_ON_ANKAN_KOKUSHI 0x01u
#define RV_RULE_KOKUSHI13_DOUBLE 0x02u
#define RV_RULE_SUUANKOU_TANKI_DOUBLE 0x04u
#define RV_RULE_JUNSEI_CHUUREN_DOUBLE 0x08u
#define RV_RULE_DAISUUSHII_DOUBLE 0x10u
#define RV_RULE_PAO_LIABILITY_ONLY 0x20u
#define RV_RULE_SANCHAHO_IS_DRAW 0x40u
#define RV_RULE_KUIKAE_FORBIDDEN 0x80u
#define RV_RULE_DEFAULT_TENHOU (RV_RULE_SANCHAHO_IS_DRAW | RV_RULE_KUIKAE_FORBIDDEN)           /* rule.rs:29-42 */
#define RV_RULE_DEFAULT_MJSOUL (0x3Fu | RV_RULE_KUIKAE_FORBIDDEN)                               /* rule.rs:44-57 */

#define RV_NONE 0xFFu /* Option::None for u8 fields */
#define RV_NP 4
#define RV_HAND_CAP 16 /* 13 + the drawn tile; the reference's hands are unbounded Vecs and its own tests park up to 15 tiles in one */
#define RV_RIVER_CAP 32
#define RV_MAX_CLAIMS 48
#define RV_MAX_LEGAL 64

/* per-seat flag bits (state/player.rs:14-34) */
#define RV_F_RIICHI_DECLARED 0x01u
#define RV_F_RIICHI_STAGE 0x02u
#define RV_F_DOUBLE_RIICHI 0x04u
#define RV_F_MISSED_AGARI_RIICHI 0x08u
#define RV_F_MISSED_AGARI_DOUJUN 0x10u
#define RV_F_NAGASHI_ELIGIBLE 0x20u
#define RV_F_IPPATSU_CYCLE 0x40u

/* ---- Action (action.rs:82-105) ------------------------------------------ */
typedef struct rv_action {
  uint8_t type;       /* rv_action_type */
  uint8_t tile;       /* tid 0..135 or RV_NONE */
  uint8_t n_consume;  /* 0..4 */
  uint8_t consume[4]; /* sorted ascending (Action::new sorts, action.rs:97-98) */
  uint8_t actor;      /* seat or RV_NONE */
} rv_action;

/* ---- per-game state: the HBM record AND the snapshot format --------------
 * Replaces GameState / PlayerState / WallState (state/mod.rs:31-91,
 * state/player.rs:6-39, state/wall.rs:9-19).  One record per game, array of
 * records in HBM.  `wall` holds the reference's `wall.tiles` Vec as a fixed
 * array: the Vec's front (rinshan side) is wall[rinshan_draw_count], its back
 * (live-draw side) is wall[wall_top-1].                                      */
typedef struct rv_game_state {
  /* ---- hot part: the first RV_HOT_BYTES bytes.  The rollout kernels stage exactly this prefix in shared
   * memory (one bulk copy per game in, one out) and reach the cold arrays below through the HBM record. ---- */
  /* derived caches, kept consistent by every mutation (the oracle recomputes them from scratch in its
   * snapshot, so state-parity tests also check the incremental maintenance):                          */
  uint64_t c_cnt[RV_NP][4];       /* concealed-hand histogram, 4-bit count per tile kind; [seat][m,p,s,z] */
  uint64_t c_river_kinds[RV_NP];  /* bit k: some own discard has kind k (furiten test)                     */
  uint64_t c_waits[RV_NP];        /* get_waits_u8 of the seat's hand when it is 13-tile-equivalent, else 0 */
  uint8_t hand[RV_NP][RV_HAND_CAP]; /* ordered as the reference's Vec (legal_actions.rs:102-110 lists discards in this order);
                                       RV_NONE pad; each row is 16-byte aligned (offset 192 + 16 seat): one 128-bit load */
  uint64_t seed;                  /* wall.seed; also the game id that keys the on-device agent */
  uint64_t ev_hash;               /* FNV-1a-64 over the 32-bit words of the binary event stream */
  uint32_t c_key[RV_NP][4];       /* base-5 suit keys of c_cnt (table indices)                             */
  uint32_t river_tedashi[RV_NP];  /* bit i = discard_from_hand[i] */
  int32_t score[RV_NP];
  uint32_t riichi_sticks;
  uint32_t turn_count;
  /* counters (not in the reference): */
  uint32_t step_count;            /* env steps taken (RiichiEnv.step calls that were not no-ops on a done game) */
  uint32_t ev_count;              /* events pushed since reset */
  uint32_t ev_words;              /* 32-bit words pushed since reset (log length, even when the log is capped/off) */

  uint8_t hand_len[RV_NP];
  uint8_t meld_tiles[RV_NP][4][4];  /* tids, order as stored by the reference (sorted); RV_NONE pad */
  uint8_t meld_type[RV_NP][4];      /* rv_meld_type */
  uint8_t n_melds[RV_NP];
  uint8_t n_river[RV_NP];
  uint8_t flags[RV_NP];             /* RV_F_* */
  uint8_t forbidden[RV_NP][2];      /* forbidden_discards (tids), RV_NONE pad */

  uint8_t wall_len;               /* 136 (4P) or 108 (3P) */
  uint8_t wall_top;               /* tiles not yet popped from the back (absolute index + 1) */
  uint8_t rinshan_draw_count;     /* wall.rs:12 */
  uint8_t pending_kan_dora_count; /* wall.rs:13 */
  uint8_t drawable_count;         /* wall.rs:14 */
  uint8_t n_dora;
  uint8_t dora_ind[5];            /* wall.dora_indicators (tids) */
  uint8_t phase;                  /* rv_phase */

  uint8_t current_player, oya, honba, kyoku_idx;
  uint8_t round_wind, is_done, needs_tsumo, is_first_turn;
  uint8_t is_rinshan_flag, riichi_pending_acceptance, drawn_tile, last_discard_pid;
  uint8_t last_discard_tile, pending_kan_pid, pending_kan_type, pending_kan_tile;
  uint8_t active_mask;            /* active_players as a seat bitmask (always ascending seat order in the reference) */
  uint8_t last_error;             /* RV_NONE or offending seat (state/mod.rs:395-399) */
  uint8_t game_mode, rule_bits;
  uint8_t overflow;               /* bit0: a fixed capacity (river / claims / hand) was exceeded; bit1: game retired on a dead end;
                                     bit2: the event log is full — later events were dropped (hash and counters go on) */
  uint8_t pending_init[3];        /* {oya, round_wind, honba} of a round whose deal is deferred inside a rollout kernel;
                                     pending_init[0]==RV_NONE outside kernels (always, as seen through this API)      */
  uint8_t n_claims[RV_NP];        /* lengths of claims[] below */
  uint8_t pending_tail[2];        /* {tile, tsumogiri} of a discard whose follow-up (_resolve_discard) is deferred inside a rollout
                                     kernel; pending_tail[0]==RV_NONE outside kernels (always, as seen through this API) */
  uint8_t is_after_kan;           /* state/mod.rs:87 — replay only (apply_log_action): the next DealTile is a rinshan draw */
  uint8_t hot_reserved[9];

  /* ---- cold part (offset RV_HOT_BYTES): large arrays touched a byte or a few words at a time ---- */
  uint8_t wall[136];
  uint8_t river[RV_NP][RV_RIVER_CAP]; /* discards */
  /* current_claims (state/mod.rs:47): packed type | tile<<8 | c0<<16 | c1<<24 */
  uint32_t claims[RV_NP][RV_MAX_CLAIMS];
  /* fields a step rarely reads: kept out of the staged prefix so that more games fit in an SM's shared memory */
  uint64_t hand_index;              /* wall.hand_index */
  uint32_t river_riichi[RV_NP];     /* bit i = discard_is_riichi[i] (written by a riichi discard, read by nobody on the step path) */
  int32_t score_delta[RV_NP];
  uint8_t meld_from[RV_NP][4];      /* from_who, RV_NONE == -1 */
  uint8_t meld_called[RV_NP][4];    /* called_tile or RV_NONE */
  uint8_t pao[RV_NP][2];            /* [seat][0]: liable seat for yaku 37, [1]: for yaku 50; RV_NONE */
  uint32_t kyoku_count;             /* rounds dealt since reset (counter, not in the reference) */
  uint8_t riichi_decl_idx[RV_NP];   /* riichi_declaration_index or RV_NONE */
  uint8_t riichi_sutehai[RV_NP];    /* state/mod.rs:89 */
  uint8_t last_tedashi[RV_NP];      /* state/mod.rs:90 */
  uint8_t n_kita[RV_NP];            /* 3P: kita count per seat */
  uint8_t reserved[4];              /* keeps sizeof a multiple of 16 (bulk-copy granularity) */
} rv_game_state;
#define RV_HOT_BYTES 544          /* == offsetof(rv_game_state, wall); a multiple of 16 */

/* ---- binary event stream (replaces _push_mjai_event, state/mod.rs:2094-2148)
 * A sequence of 32-bit words.  word0 = type | nwords<<8 | a<<16 | b<<24.
 * Host code (rv_event_to_json) renders MJAI JSON with alphabetical keys.     */
enum rv_event_type {
  RV_EV_START_GAME = 1,   /* 1 word */
  RV_EV_START_KYOKU = 2,  /* a=bakaze(round_wind%4) b=oya; w1=honba|dora_marker<<8|kyotaku<<16; w2..5=scores; w6..18=tehais 4x13 bytes */
  RV_EV_TSUMO = 3,        /* a=actor b=tile */
  RV_EV_DAHAI = 4,        /* a=actor b=tile; tsumogiri=false */
  RV_EV_DAHAI_TSUMOGIRI = 5,
  RV_EV_REACH = 6,        /* a=actor */
  RV_EV_REACH_ACCEPTED = 7,
  RV_EV_PON = 8,          /* a=actor b=tile; w1=target|c0<<8|c1<<16|0xFF<<24 */
  RV_EV_CHI = 9,
  RV_EV_DAIMINKAN = 10,   /* w1=target|c0<<8|c1<<16|c2<<24 */
  RV_EV_ANKAN = 11,       /* a=actor b=pai; w1=c0..c3 */
  RV_EV_KAKAN = 12,       /* a=actor b=pai; w1=c0|c1<<8|c2<<16|0xFF<<24 */
  RV_EV_DORA = 13,        /* b=dora_marker */
  RV_EV_HORA = 14,        /* a=actor b=target; w1=tsumo|n_ura<<8|han<<16|fu<<24; w2=ura0..3; w3=ura4|yakuman<<8; w4..7=deltas; w8,w9=yaku bitmask lo,hi */
  RV_EV_RYUKYOKU = 15,    /* a=reason code; w1..4=deltas */
  RV_EV_END_KYOKU = 16,
  RV_EV_END_GAME = 17,
  RV_EV_KITA = 18         /* a=actor b=tile */
};
enum rv_ryukyoku_reason {
  RV_RK_EXHAUSTIVE = 0, RV_RK_NAGASHI = 1, RV_RK_KYUSHU = 2, RV_RK_SUFUURENTA = 3,
  RV_RK_SUUKANSANSEN = 4, RV_RK_SUUCHA_RIICHI = 5, RV_RK_SANCHAHO = 6,
  RV_RK_ILLEGAL_BASE = 8 /* + offending seat: "Error: Illegal Action by Player N" */
};

/* ---- hand evaluation (config 2) ------------------------------------------
 * Replaces HandEvaluator::{new,calc,get_waits_u8,is_tenpai}
 * (hand_evaluator.rs:24-213), agari::is_agari (agari.rs:63),
 * yaku::calculate_yaku (yaku.rs:232), score::calculate_score (score.rs:13),
 * shanten::calculate_shanten (shanten.rs:250).                              */
#define RV_C_TSUMO 0x001u
#define RV_C_RIICHI 0x002u
#define RV_C_DOUBLE_RIICHI 0x004u
#define RV_C_IPPATSU 0x008u
#define RV_C_HAITEI 0x010u
#define RV_C_HOUTEI 0x020u
#define RV_C_RINSHAN 0x040u
#define RV_C_CHANKAN 0x080u
#define RV_C_TSUMO_FIRST_TURN 0x100u

typedef struct rv_hand_query {  /* 56 bytes */
  uint8_t tiles[14];       /* concealed tids, RV_NONE pad */
  uint8_t n_tiles;
  uint8_t n_melds;
  uint8_t meld_type[4];    /* rv_meld_type */
  uint8_t meld_tiles[4][4];/* tids (RV_NONE pad for 3-tile melds) */
  uint8_t win_tile;        /* tid */
  uint8_t n_dora, n_ura;
  uint8_t dora_ind[5];     /* tids */
  uint8_t ura_ind[5];
  uint8_t player_wind, round_wind; /* 0..3 */
  uint8_t honba;
  uint16_t cond;           /* RV_C_* */
  uint8_t sanma;           /* 1: HandEvaluator3P semantics (hand_evaluator_3p.rs: 1m<->9m dora wrap, nukidora, 3-player score) */
  uint8_t kita_count;      /* Conditions.kita_count (3P) */
} rv_hand_query;

typedef struct rv_hand_result { /* 40 bytes */
  uint64_t yaku_mask;      /* bit id set for every yaku id in WinResult.yaku */
  uint64_t wait_mask;      /* get_waits_u8 of the 3n+1 hand (tiles without win tile when 3n+2), bit t34 */
  uint32_t ron_agari, tsumo_agari_oya, tsumo_agari_ko;
  uint8_t is_win, yakuman, has_win_shape, han;
  uint8_t fu;
  int8_t shanten;          /* calculate_shanten(tiles [+ win tile when 3n+1]) */
  int8_t shanten13;        /* calculate_shanten of the 3n+1 hand */
  uint8_t n_yaku;
  uint8_t _pad[4];
} rv_hand_result;

typedef struct rv_ctx rv_ctx; /* opaque: device tables + stream */

const char* rv_last_error(void);
int rv_version(void);

/* Create / destroy a context on CUDA device `device` (uploads lookup tables). */
int rv_ctx_create(int device, rv_ctx** out);
int rv_ctx_destroy(rv_ctx* ctx);
int rv_ctx_sync(rv_ctx* ctx);
/* The context's cudaStream_t (as void*), so callers can order their own work (e.g. torch.cuda.ExternalStream). */
void* rv_ctx_stream(rv_ctx* ctx);
/* CUDA-event timers on the context stream: mark records event idx; elapsed synchronises on idx_b and returns ms. */
int rv_timer_mark(rv_ctx* ctx, int idx /*0..7*/);
int rv_timer_elapsed(rv_ctx* ctx, int idx_a, int idx_b, float* ms);
/* sizeof() of the public structs as compiled: 0 rv_game_state, 1 rv_hand_query, 2 rv_hand_result, 3 rv_action,
 * 4 rv_mjai_event, 5 rv_run_stats, 6 rv_log_action, 7 rv_log_kyoku */
int rv_sizeof(int which);

/* Batched hand evaluation with HOST buffers (copies inside the call). */
int rv_hand_eval_batch(rv_ctx* ctx, const rv_hand_query* q, rv_hand_result* out, int64_t n);
/* Same with DEVICE buffers; asynchronous on the context's stream. */
int rv_hand_eval_batch_device(rv_ctx* ctx, const rv_hand_query* d_q, rv_hand_result* d_out, int64_t n);
/* The seeded synthetic hand stream of BASELINE.json configs[1] (definition: include/rv_synth.h): queries first .. first+n-1
 * written into a DEVICE buffer, asynchronous on the context's stream; the _host variant builds one query on the CPU. */
int rv_hand_queries_seeded(rv_ctx* ctx, rv_hand_query* d_q, uint64_t first, int64_t n);
int rv_hand_query_seeded_host(uint64_t h, rv_hand_query* out);
/* score.rs:13-52 on host (no device work): out[0]=pay_ron out[1]=pay_tsumo_oya out[2]=pay_tsumo_ko out[3]=total */
int rv_calculate_score(int han, int fu, int is_oya, int is_tsumo, uint32_t honba, int num_players, uint32_t out[4]);

/* ---- the vectorised environment -------------------------------------------
 * N independent games, SoA-of-records in HBM.  Replaces a Python list of
 * RiichiEnv objects (riichienv-ml/.../_ppo_worker.py:39).                    */
typedef struct rv_vec rv_vec; /* opaque */

/* RiichiEnv::new (env.rs:82-118) for n games: game g gets seed seeds[g]
 * (seeds==NULL -> seed_base+g).  The constructor's shuffle #0 is consumed as
 * in the reference (state/mod.rs:165).  log_cap_words: per-game capacity of
 * the binary event log in 32-bit words, 0 == skip_mjai_logging=True (only the
 * rolling hash is kept).                                                     */
int rv_vec_create(rv_ctx* ctx, int64_t n, int game_mode, uint32_t rule_bits, const uint64_t* seeds,
                  uint64_t seed_base, uint32_t log_cap_words, rv_vec** out);
int rv_vec_destroy(rv_vec* v);
/* Re-seed every game as if freshly constructed with RiichiEnv(seed=...) (hand_index back to 1); asynchronous.
 * seeds: HOST array of n u64 (copied H2D) or NULL -> seed_base + g.  Follow with rv_vec_reset.            */
int rv_vec_reseed(rv_vec* v, const uint64_t* seeds, uint64_t seed_base);
int64_t rv_vec_size(const rv_vec* v);

/* RiichiEnv::reset (env.rs:799-851) for every game.  Optional per-game arrays
 * (NULL = default): oya[n], round_wind[n], honba[n], kyotaku[n], scores[n*NP],
 * walls[n*wall_len] (reset(wall=) -> load_wall, wall.rs:69-80).              */
int rv_vec_reset(rv_vec* v, const uint8_t* oya, const uint8_t* round_wind, const uint8_t* honba,
                 const uint32_t* kyotaku, const int32_t* scores, const uint8_t* walls);

/* Legal actions of every seat that owes an action (Observation.legal_actions,
 * observation/mod.rs:113-115; order of state/legal_actions.rs:11-508).
 * out_actions: [n][NP][RV_MAX_LEGAL], out_counts: [n][NP] (0 for idle seats). */
int rv_vec_legal_actions(rv_vec* v, rv_action* out_actions, uint8_t* out_counts);

/* RiichiEnv::step (env.rs:857-872 -> state/mod.rs:330-1315).  actions:
 * [n][NP]; type RV_NO_ACTION == key absent.  Illegal actions trigger the
 * reference's penalty ryukyoku and set last_error.                           */
int rv_vec_step(rv_vec* v, const rv_action* actions);

/* Persistent on-device rollout with the keyed random agent (SURVEY.md §8 d):
 * every game takes up to max_steps env steps (stopping at done):
 *   idx = mix64(agent_seed ^ game_id*0x9E3779B97F4A7C15 ^ (step_count<<8) ^ seat) % n_legal.
 * Returns the number of env steps executed by all games in *steps_done.      */
int rv_vec_step_random(rv_vec* v, uint64_t agent_seed, uint32_t max_steps, uint64_t* steps_done);
/* Asynchronous variant on the context stream: accumulates into a device counter; read with rv_vec_steps_total. */
int rv_vec_step_random_async(rv_vec* v, uint64_t agent_seed, uint32_t max_steps);
int rv_vec_steps_total(rv_vec* v, uint64_t* steps_total, int64_t* games_done);

/* The same loop with a selectable on-device agent (README.md:50-62 with `agent.act` on the device):
 *   RV_AGENT_RANDOM  the uniform keyed agent of rv_vec_step_random (src/riichienv/agents/random_agent.py:6-15)
 *   RV_AGENT_GREEDY  a keyed "greedy-win" agent: every Tsumo / Ron, every Riichi, Pon / Kan / Kita with probability 1/4, Chi
 *                    with 1/8, otherwise the discard that leaves the lowest shanten (ties keyed).  Its rollouts end ~60 % of
 *                    the rounds in a win; it exists so that parity runs exercise the settlement path (state/mod.rs:685-893,
 *                    919-1142).  Thread-per-game kernel: an evaluation path, not the throughput path.                       */
enum rv_agent_policy { RV_AGENT_RANDOM = 0, RV_AGENT_GREEDY = 1 };
int rv_vec_step_agent(rv_vec* v, int policy, uint64_t agent_seed, uint32_t max_steps, uint64_t* steps_done);

/* RiichiEnv::{done,scores,ranks} (env.rs:353,401,673-689).  Host outputs:
 * done[n], scores[n*NP], ranks[n*NP] (1-based), any may be NULL.            */
int rv_vec_results(rv_vec* v, uint8_t* done, int32_t* scores, uint8_t* ranks);
/* per-game counters: step_count[n], kyoku_count[n], ev_count[n], ev_hash[n] (any may be NULL) */
int rv_vec_counters(rv_vec* v, uint32_t* step_count, uint32_t* kyoku_count, uint32_t* ev_count, uint64_t* ev_hash);

/* Snapshot get/set (RiichiEnv getters/setters, env.rs:134-635; clone() 358-372). */
int rv_vec_get_state(rv_vec* v, int64_t game, rv_game_state* out);
int rv_vec_set_state(rv_vec* v, int64_t game, const rv_game_state* in);
/* RiichiEnv::clone / __copy__ / __deepcopy__ (env.rs:358-372): an independent vector with the same games — records,
 * event logs and sequence-feature cursors copied device to device.                                             */
int rv_vec_clone(rv_vec* v, rv_vec** out);
/* RiichiEnv::_reveal_kan_dora (op 0; *n_out = number of dora indicators afterwards) and RiichiEnv::_get_ura_markers
 * (op 1; out_tiles[5] = ura indicator tile ids, RV_NONE pad, *n_out = how many) of env.rs:624-631: the reference exposes
 * these two internals to its tests (tests/env/test_paishan.py); they run the device routines the step path uses.
 * Ops 2-5 are what the reference's Rust unit tests call directly on a GameState (riichienv-core/src/tests.rs:172-262,
 * 375-428): 2 = _trigger_ryukyoku("exhaustive_draw"); 3 / 4 / 5 = _initialize_next_round(false,false) / (true,false) /
 * (false,true); *n_out = is_done afterwards.  Op 6 is the replay iterator's `_get_claim_actions_for_player` for every seat
 * against the game's last discard (replay/mod.rs:130-181): claim lists, active seats and phase are left in the record
 * (the caller reads them with rv_vec_get_state and puts the record back); *n_out = number of seats with a claim.        */
int rv_vec_debug_call(rv_vec* v, int64_t game, int op, uint8_t out_tiles[5], int* n_out);
/* Device pointer to the state records (for zero-copy consumers). */
int rv_vec_state_device_ptr(rv_vec* v, void** d_states);

/* Binary event log of one game (mjai_log, env.rs:729-739).  Copies up to cap words; *n_words receives the readable
 * length: the words the log holds, cut back to the last whole event if the per-game capacity was exceeded (the record's
 * overflow bit 2 says so; rv_vec_counters' ev_count / ev_hash still cover every event).                                */
int rv_vec_events(rv_vec* v, int64_t game, uint32_t* out_words, uint32_t cap, uint32_t* n_words);
/* Render one binary event as MJAI JSON (alphabetical keys, as serde_json
 * without preserve_order).  viewer = seat for the per-player masked log
 * (state/mod.rs:2109-2143) or -1 for the global log.  Returns words consumed. */
int rv_event_to_json(const uint32_t* words, uint32_t n_words, int viewer, char* out, uint32_t out_cap);

/* Seeded wall (state/wall.rs:36-67): host implementation for tests / tools. */
int rv_wall_from_seed(uint64_t seed, uint64_t hand_index, int n_tiles /*136|108*/, uint8_t* out_tiles_reversed);

/* Observation tensors for every seat that owes an action (what RiichiEnv.step returns observations for):
 * Observation::encode (observation/python.rs:457-806) and Observation::mask (98-111).  Rows are written
 * compactly in ascending (game, seat) order:
 *   d_obs   [max_obs][74][34] f32   (device, 16-byte aligned)      — may be NULL
 *   d_mask  [max_obs][82] u8        (device)                       — may be NULL
 *           sanma (Observation3P::encode / mask, observation_3p/python.rs:402-708, 102-114): [74][27] f32 and [60] u8
 *   d_index [max_obs] i32           (device) game*4 + seat per row — may be NULL
 * *n_obs (host, may be NULL -> fully asynchronous) receives the number of rows; rows beyond max_obs are dropped. */
int rv_vec_encode(rv_vec* v, float* d_obs, uint8_t* d_mask, int32_t* d_index, int64_t max_obs, int64_t* n_obs);

/* Extended observation tensors: Observation::encode_extended (observation/python.rs:1272-1294; channel blocks
 * observation/encode.rs:12-584) — base 74 channels + discard decay, shanten efficiency (shanten.rs:250-393), ankan / fuuro
 * overview, action availability, discard candidates, pass context, last tedashi, riichi sutehai = 215 channels.
 * Same row order and arguments as rv_vec_encode:
 *   d_obs   [max_obs][215][34] f32  (device, 8-byte aligned)       — may be NULL
 *   d_mask  [max_obs][82] u8, d_index [max_obs] i32                — may be NULL
 *   sanma (Observation3P::encode_extended, observation_3p/python.rs:1117-1140): [215][27] f32 (4-byte aligned) and [60] u8 */
int rv_vec_encode_ext(rv_vec* v, float* d_obs, uint8_t* d_mask, int32_t* d_index, int64_t max_obs, int64_t* n_obs);

/* Observation::encode_kawa_overview (observation/python.rs:881-930) for every seat that owes an action, same row order as
 * rv_vec_encode: d_out [max_obs][4][7][34] f32 (device, 16-byte aligned), seats in absolute order;
 * sanma (Observation3P::encode_kawa_overview, observation_3p/python.rs:760-808): [max_obs][3][7][27] f32 (4-byte aligned). */
int rv_vec_encode_kawa(rv_vec* v, float* d_out, int32_t* d_index, int64_t max_obs, int64_t* n_obs);

/* Observe + step in one pass (BASELINE config 5: a random-agent rollout that emits FEATURE_ENCODING tensors at every step —
 * the loop `obs = env.step({p: agent.act(o) ...})` of README.md:50-62 with the agent of rv_vec_step_random on the device).
 * Exactly rv_vec_encode(v, d_obs, d_mask, d_index, max_obs, n_obs) followed by rv_vec_step_random_async(v, agent_seed, 1):
 * the rows describe the decision point BEFORE the step, then every live game takes one env step.  One kernel reads each
 * game record once: the tensors are written from the staged state and the masks from the legal lists the agent picks from. */
int rv_vec_observe_step_random(rv_vec* v, uint64_t agent_seed, float* d_obs, uint8_t* d_mask, int32_t* d_index, int64_t max_obs,
                               int64_t* n_obs);

/* Sequence features (observation/sequence_features.rs:331-835; docs/SEQUENCE_FEATURE_ENCODING.md) for every seat that
 * owes an action, rows in the same (game, seat) order as rv_vec_encode.  4P only (as in the reference); the vector
 * must have been created with an event log (log_cap_words > 0).
 *   d_sparse  [max_obs][25] u16     encode_seq_sparse(game_style), padded with 441            — may be NULL
 *   d_numeric [max_obs][12] f32     encode_seq_numeric                                       — may be NULL
 *   d_prog    [max_obs][max_prog][5] u16  encode_seq_progression, padded with (4,276,2,2,4)   — may be NULL
 *   d_cand    [max_obs][64][4] u16  encode_seq_candidates, padded with (279,2,2,3)            — may be NULL
 *   d_lens    [max_obs][3] u16      lengths the reference would return: sparse, progression (<= 512), candidates
 *   d_index   [max_obs] i32         game*4 + seat                                             — may be NULL
 * The features are functions of the seat's EVENT DELTA: the events pushed since that seat's previous observation
 * (state/mod.rs:211-218; the live env never enables the progression cache, state/mod.rs:146).
 *   start_words == NULL: the library keeps one cursor per (game, seat) — zeroed by rv_vec_reset, advanced to the end
 *                        of the log by every rv_vec_encode_seq call for the seats it observed;
 *   start_words != NULL: host array [n][4] of word offsets into each game's event log; internal cursors untouched. */
int rv_vec_encode_seq(rv_vec* v, int game_style, const uint32_t* start_words, uint16_t* d_sparse, float* d_numeric,
                      uint16_t* d_prog, int max_prog, uint16_t* d_cand, uint16_t* d_lens, int32_t* d_index, int64_t max_obs,
                      int64_t* n_obs);

/* ---- MJAI-driven state tracking (RiichiEnv::apply_event / observe_event, env.rs:880-948) -------------------------------
 * One parsed MJAI event; tiles are ids 0..135 — mjai_to_tid (parser.rs:336-385) on the host, an unparsable tile ("?") is 0
 * as in event_handler.rs:8-10.  `type` is an rv_event_type; RV_EV_NONE (0) = no event for this game / an event type the
 * reference ignores.  Fields beyond those an event type carries are ignored.                                              */
#define RV_EV_NONE 0
typedef struct rv_mjai_event {
  uint8_t type;          /* rv_event_type; daiminkan ("kan") = RV_EV_DAIMINKAN; dahai with tsumogiri = RV_EV_DAHAI_TSUMOGIRI */
  uint8_t actor, target;
  uint8_t pai;           /* tile id (dora: the dora_marker) */
  uint8_t n_consumed;
  uint8_t consumed[4];
  uint8_t bakaze;        /* start_kyoku: 0..3 = E S W N */
  uint8_t kyoku, honba, oya, dora_marker;
  uint8_t tehai_len[RV_NP];
  uint8_t tehais[RV_NP][14];
  uint8_t _pad[3];
  uint32_t kyotaku;
  int32_t scores[RV_NP];
} rv_mjai_event;
/* GameState::apply_mjai_event (state/event_handler.rs:18-330; sanma state_3p/event_handler.rs:19-362) for every game:
 * events[n], type RV_EV_NONE = leave that game alone.  A start_game event also resets the game's logs (env.rs:56-72).
 * The event is NOT appended to the device event log: the caller keeps the text it fed (RiichiEnv.apply_event does).     */
int rv_vec_apply_events(rv_vec* v, const rv_mjai_event* events);

/* ---- replay ingestion: MJAI logs -> kyoku records + log actions -> batched state tracking ----------------
 * Replaces MjaiReplay::from_jsonl / KyokuBuilder (riichienv-core/src/replay/mjai_replay.rs:184-633), the replay `Action`
 * enum (replay/mod.rs:35-96), the state set-up of LogKyoku::steps (replay/mod.rs:1094-1292) and GameState::apply_log_action
 * (state/event_handler.rs:332-894; sanma state_3p/event_handler.rs:365-808).  One game record follows one kyoku: a batch of
 * N kyoku is replayed in lock-step, one log action per record per call, and the decision points in between are read with
 * the ordinary observation calls (rv_vec_legal_actions, rv_vec_encode, ...).                                                */
enum rv_log_action_type {
  RV_LA_NONE = 0,            /* no action for this record / Action::Other */
  RV_LA_DISCARD = 1,         /* DiscardTile  {seat, tile, flags bit0 is_liqi, bit1 is_wliqi} */
  RV_LA_DEAL = 2,            /* DealTile     {seat, tile} */
  RV_LA_CHI_PENG_GANG = 3,   /* ChiPengGang  {seat, meld_type, tiles[n_tiles], froms[n_tiles]}: tiles[0] = the called tile */
  RV_LA_ANGANG_ADDGANG = 4,  /* AnGangAddGang{seat, meld_type (RV_MELD_ANKAN | RV_MELD_KAKAN), tiles[n_tiles]} */
  RV_LA_DORA = 5,            /* Dora         {tile = dora_marker} */
  RV_LA_HULE = 6,            /* Hule         {hules[n_hule]} */
  RV_LA_NOTILE = 7,          /* NoTile (exhaustive draw) */
  RV_LA_BABEI = 8,           /* BaBei        {seat, flags bit0 moqie} (sanma Kita) */
  RV_LA_LIUJU = 9            /* LiuJu        {seat, tile = lj_type} (abortive draw) */
};
typedef struct rv_hule {       /* HuleData, replay/mod.rs:82-96 */
  uint8_t seat, hu_tile, zimo, yiman;
  uint8_t n_li_doras;          /* 0xFF = li_doras is None */
  uint8_t li_doras[5];
  uint8_t _pad[2];
  uint32_t count, fu;          /* han (or yakuman count), fu */
  uint32_t point_rong, point_zimo_qin, point_zimo_xian;
  uint64_t fans;               /* bit y = yaku id y is in `fans` */
} rv_hule;
typedef struct rv_log_action {
  uint8_t type;                /* rv_log_action_type */
  uint8_t seat, tile, flags;
  uint8_t meld_type, n_tiles;
  uint8_t tiles[4], froms[4];
  uint8_t n_hule;
  uint8_t _pad;
  rv_hule hules[3];
} rv_log_action;
#define RV_LOG_MAX_DORAS 8
typedef struct rv_log_kyoku {  /* LogKyoku, replay/mod.rs:1010-1029 (+ what LogKyoku::steps derives before the first action) */
  uint8_t np;                  /* scores.len(): 4, or 3 for sanma */
  uint8_t chang, ju, ben, liqibang, left_tile_count;
  uint8_t n_doras, n_ura_doras;
  uint8_t doras[RV_LOG_MAX_DORAS], ura_doras[RV_LOG_MAX_DORAS];
  uint8_t hand_len[RV_NP];
  uint8_t hands[RV_NP][14];
  uint8_t wliqi[RV_NP];
  uint8_t oya;                 /* steps(): ju % np, or the seat dealt 14 tiles (replay/mod.rs:1125-1132) */
  uint8_t oya_drawn_tile;      /* steps(): drawn tile of a 14-tile dealer read off the first action, RV_NONE for 13-tile deals */
  uint8_t has_game_end_scores;
  uint8_t _pad[3];
  uint32_t rule_bits;
  int32_t n_actions;
  int32_t scores[RV_NP], end_scores[RV_NP], game_end_scores[RV_NP];
} rv_log_kyoku;
typedef struct rv_replay rv_replay; /* opaque: the rounds of one parsed log (host memory only) */
/* MjaiReplay::from_jsonl (mjai_replay.rs:276-370): JSON lines, gzip detected by its magic bytes.  rule_bits: RV_RULE_* preset
 * stored with every kyoku.  RV_ERR_INVALID with rv_last_error() = "Failed to open file…" / "Parse error…" as the reference raises. */
int rv_replay_from_jsonl(const char* path, uint32_t rule_bits, rv_replay** out);
/* the same parser over text already in memory (len bytes of JSON lines, not compressed) */
int rv_replay_from_text(const char* text, size_t len, uint32_t rule_bits, rv_replay** out);
/* MjSoulReplay::from_json / from_dict (replay/mjsoul_replay.rs:174-343): a paifu as JSON — {"rounds": [[{"name", "data"}, ...]]},
 * {"header", "data": [[...]]} or the bare list of rounds; _json reads a (gzip) file, _text takes the JSON text.  The first
 * action of every round (NewRound) is kept as an RV_LA_NONE placeholder, as the reference keeps Action::Other.             */
int rv_replay_from_mjsoul_json(const char* path, uint32_t rule_bits, rv_replay** out);
int rv_replay_from_mjsoul_text(const char* text, size_t len, uint32_t rule_bits, rv_replay** out);
/* Bulk loading for a data loader (riichienv-ml/.../datasets/mjai_logs.py:73-84 opens the files one by one): n_paths files —
 * format 0 = MJAI JSON lines (rv_replay_from_jsonl), 1 = MjSoul paifu (rv_replay_from_mjsoul_json) — parsed by `threads` host
 * threads (<= 0: one per core) into ONE replay whose rounds follow the order of `paths`.  A file that does not open or parse is
 * skipped and counted in *n_failed, as the reference's datasets skip it.                                                   */
int rv_replay_from_files(const char* const* paths, int n_paths, int format, uint32_t rule_bits, int threads, rv_replay** out,
                         int* n_failed);
/* rounds with np seats and their actions in total; rv_replay_flatten writes exactly these: kyokus[n_rounds], the action lists
 * back to back in actions[n_actions], first[n_rounds + 1] offsets (the arguments of rv_vec_replay_load) and, if not NULL, the
 * index of each in the replay (round_index[n_rounds]).                                                                     */
int rv_replay_totals(const rv_replay* r, int np, int64_t* n_rounds, int64_t* n_actions);
int rv_replay_flatten(const rv_replay* r, int np, rv_log_kyoku* kyokus, rv_log_action* actions, int64_t* first, int32_t* round_index);
/* What the seat on turn decided, per log action: seat[i] / action_id[i] (Action::encode ids, action.rs:158-227; sanma ids for
 * np = 3) when action i is a discard (a riichi discard: the Riichi id), an ankan / kakan, a tsumo win or a kita — the samples
 * Kyoku.steps yields for the seat on turn — else -1 / -1 (draws, calls and ron are not own-turn decisions).                 */
int rv_replay_own_turn_labels(const rv_log_action* actions, int64_t n, int np, int16_t* seat, int16_t* action_id);
int rv_replay_free(rv_replay* r);
int rv_replay_num_rounds(const rv_replay* r);                                  /* MjaiReplay::num_rounds */
int rv_replay_kyoku(const rv_replay* r, int round, rv_log_kyoku* out);
/* copies min(cap, n_actions) actions of the round; *n_out = n_actions */
int rv_replay_actions(const rv_replay* r, int round, rv_log_action* out, int cap, int* n_out);
/* The Option fields of the reference's Action that rv_log_action does not carry (state tracking never reads them; Kyoku.events()
 * and the win-context walk do): one record per action of the round, as rv_replay_actions.  All "none" for MJAI logs.        */
typedef struct rv_log_action_aux {
  uint8_t n_doras;                 /* DiscardTile / DealTile `doras`; 0xFF = None */
  uint8_t doras[RV_LOG_MAX_DORAS];
  uint8_t left_tile_count;         /* DealTile `left_tile_count`; 0xFF = None */
  uint8_t tile_raw_id;             /* AnGangAddGang `tile_raw_id` (kind 0..33; 0 in MJAI logs as in the reference) */
  uint8_t _pad;
} rv_log_action_aux;
int rv_replay_actions_aux(const rv_replay* r, int round, rv_log_action_aux* out, int cap, int* n_out);
/* LogKyoku.paishan (replay/mod.rs:1022,1077): the wall string a MjSoul NewRound record carries; copies min(cap, len) bytes (no
 * terminator), *n_out = len, or -1 when the log has none (every MJAI log). */
int rv_replay_paishan(const rv_replay* r, int round, char* out, int cap, int* n_out);
/* WinResultContextIterator (replay/mod.rs:1594-2093): the walk over a kyoku that tracks hands, melds, dora markers and the win
 * conditions up to every Hule.  One record per winner, in log order; `query` is what the iterator hands HandEvaluator::calc, so
 * WinResultContext.actual of a whole log is ONE rv_hand_eval_batch over the records' queries (MjSoulReplay::verify,
 * mjsoul_replay.rs:357-430, compares it with the expected_* fields the paifu recorded).  As the reference: the 4-player
 * evaluator also for sanma logs, honba and riichi sticks 0.  round < 0 = every round of the log; copies min(cap, n) records,
 * *n_out = n.  RV_ERR_INVALID for logs the reference's walk would panic on (a seat the kyoku lacks, more than 14 tiles).   */
typedef struct rv_win_context {
  rv_hand_query query;
  uint64_t expected_yaku;      /* bit id = id in HuleData.fans (paifu fans with val > 0; none in MJAI logs) */
  uint32_t expected_han, expected_fu;
  int32_t round, action;       /* the kyoku and the index of the Hule action in it */
  uint8_t seat;
  int8_t meld_from[4];         /* Meld.from_who of query.meld_*[m] (-1: none) */
  uint8_t meld_called[4];      /* Meld.called_tile (RV_NONE: none) */
  uint8_t _pad[7];
} rv_win_context;
int rv_replay_win_contexts(const rv_replay* r, int round, rv_win_context* out, int cap, int* n_out);
int rv_replay_sizeof(int which); /* sizeof() as compiled: 0 rv_win_context, 1 rv_log_action_aux (a binding's layout check, like rv_sizeof) */
/* encode_seq_progression of a REPLAY observation: in replay mode GameState::apply_log_action keeps a progression cache
 * (state/event_handler.rs:351-379, 443-486, 573-606 -> observation/sequence_features.rs:212-313) and the observation returns it
 * as it stands (sequence_features.rs:505-507).  The tuples (actor, type, moqie, liqi, from) of actions[0..n); tsumogiri[i] != 0
 * = the discard of action i was the tile just drawn (the state's drawn_tile at that point).  Writes min(cap, m) rows of 5
 * uint16, *n_out = m.  Pure host arithmetic on the log.                                                                  */
int rv_replay_progression(const rv_log_action* actions, const uint8_t* tsumogiri, int n, uint16_t* out, int cap, int* n_out);
/* LogKyoku::steps' state set-up for every game of the vector: kyokus[n] (HOST).  Each game is re-initialised as
 * `_initialize_round(oya, chang, ben, liqibang, None, scores)` does and then patched with the logged hands, dora markers and the
 * dealer's draw.  The vector's game mode must have kyoku.np seats.                                                          */
int rv_vec_replay_begin(rv_vec* v, const rv_log_kyoku* kyokus);
/* GameState::apply_log_action for every game: actions[n] (HOST), type RV_LA_NONE = leave that game alone. */
int rv_vec_apply_log_actions(rv_vec* v, const rv_log_action* actions);
/* The same with the logs resident in HBM: rv_vec_replay_load = rv_vec_replay_begin + ONE upload of every record's action
 * list (actions of record i = actions[first[i] .. first[i + 1]), first[n + 1] HOST offsets); each rv_vec_replay_advance then
 * applies the next action of every record that has one left (nothing crosses PCIe), *n_applied = how many records did
 * (0 = every kyoku is exhausted; NULL = do not wait for the kernel).                                                     */
int rv_vec_replay_load(rv_vec* v, const rv_log_kyoku* kyokus, const rv_log_action* actions, const int64_t* first);
int rv_vec_replay_advance(rv_vec* v, int64_t* n_applied);

/* ---- several GPUs behind one handle (SURVEY.md §8 e) -------------------------------------------------------------
 * Replaces what the reference does with one Python list of RiichiEnv per Ray actor (riichienv-ml/.../_ppo_worker.py:13,39).
 * Device k of G owns the contiguous global game ids [k*N/G, (k+1)*N/G); game g is seeded seed_base + g for any G, so scores,
 * ranks and event hashes do not depend on the number of devices.  Every call runs one host thread per device, each on its
 * device's own context and stream; no data crosses devices on the step path.  rv_multi_stats sums the per-device episode
 * statistics on the host (the only reduction).  Host output arrays are in global game order.                              */
typedef struct rv_multi rv_multi; /* opaque */
typedef struct rv_run_stats {
  int64_t games, games_done, env_steps, rounds;
  int64_t score_sum[RV_NP];          /* sum of final scores per seat */
  int64_t rank_hist[RV_NP][RV_NP];   /* [seat][rank - 1] over finished games */
} rv_run_stats;
int rv_multi_create(const int* devices, int n_devices, int64_t n_games, int game_mode, uint32_t rule_bits, uint64_t seed_base,
                    uint32_t log_cap_words, rv_multi** out);
int rv_multi_destroy(rv_multi* m);
int rv_multi_devices(const rv_multi* m);
int64_t rv_multi_size(const rv_multi* m);
/* the vector of shard k (for the per-device calls: encoders, snapshots) and its range of global game ids */
int rv_multi_shard(rv_multi* m, int k, rv_vec** vec, int64_t* first_game, int64_t* n_games);
int rv_multi_reset(rv_multi* m);                           /* RiichiEnv::reset() defaults for every game */
int rv_multi_reseed(rv_multi* m, uint64_t seed_base);      /* as rv_vec_reseed with seed seed_base + global id */
int rv_multi_step_random(rv_multi* m, uint64_t agent_seed, uint32_t max_steps, uint64_t* steps_done);
int rv_multi_results(rv_multi* m, uint8_t* done, int32_t* scores, uint8_t* ranks);
int rv_multi_counters(rv_multi* m, uint32_t* step_count, uint32_t* kyoku_count, uint32_t* ev_count, uint64_t* ev_hash);
int rv_multi_stats(rv_multi* m, rv_run_stats* out);

#ifdef __cplusplus
}
#endif
#endif
