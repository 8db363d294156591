"""ctypes mirror of include/riichienv_b200.h (structs + constants).

Host-side plumbing only: the layouts must match the C header byte for byte
(tests/test_abi.py checks sizeof() against the library).
"""
import ctypes as C

NP = 4
HAND_CAP = 16
RIVER_CAP = 32
MAX_CLAIMS = 48
MAX_LEGAL = 64
NONE = 0xFF

# action.rs:55-68
DISCARD, CHI, PON, DAIMINKAN, RON, RIICHI, TSUMO, PASS, ANKAN, KAKAN, KYUSHU_KYUHAI, KITA = range(12)
NO_ACTION = 255
# rule.rs:10-20
RULE_RON_ON_ANKAN_KOKUSHI = 0x01
RULE_KOKUSHI13_DOUBLE = 0x02
RULE_SUUANKOU_TANKI_DOUBLE = 0x04
RULE_JUNSEI_CHUUREN_DOUBLE = 0x08
RULE_DAISUUSHII_DOUBLE = 0x10
RULE_PAO_LIABILITY_ONLY = 0x20
RULE_SANCHAHO_IS_DRAW = 0x40
RULE_KUIKAE_FORBIDDEN = 0x80
RULE_DEFAULT_TENHOU = RULE_SANCHAHO_IS_DRAW | RULE_KUIKAE_FORBIDDEN
RULE_DEFAULT_MJSOUL = 0x3F | RULE_KUIKAE_FORBIDDEN

F_RIICHI_DECLARED = 0x01
F_RIICHI_STAGE = 0x02
F_DOUBLE_RIICHI = 0x04
F_MISSED_AGARI_RIICHI = 0x08
F_MISSED_AGARI_DOUJUN = 0x10
F_NAGASHI_ELIGIBLE = 0x20
F_IPPATSU_CYCLE = 0x40

C_TSUMO, C_RIICHI, C_DOUBLE_RIICHI, C_IPPATSU, C_HAITEI, C_HOUTEI, C_RINSHAN, C_CHANKAN, C_TSUMO_FIRST_TURN = (
    0x001, 0x002, 0x004, 0x008, 0x010, 0x020, 0x040, 0x080, 0x100)

(EV_START_GAME, EV_START_KYOKU, EV_TSUMO, EV_DAHAI, EV_DAHAI_TSUMOGIRI, EV_REACH, EV_REACH_ACCEPTED, EV_PON, EV_CHI,
 EV_DAIMINKAN, EV_ANKAN, EV_KAKAN, EV_DORA, EV_HORA, EV_RYUKYOKU, EV_END_KYOKU, EV_END_GAME, EV_KITA) = range(1, 19)

u8, u32, i32, u64, u16, i8 = C.c_uint8, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint16, C.c_int8


class Action(C.Structure):
    _fields_ = [("type", u8), ("tile", u8), ("n_consume", u8), ("consume", u8 * 4), ("actor", u8)]


class GameState(C.Structure):
    _fields_ = [
        # hot part (first HOT_BYTES bytes)
        ("c_cnt", (u64 * 4) * NP), ("c_river_kinds", u64 * NP), ("c_waits", u64 * NP),
        ("hand", (u8 * HAND_CAP) * NP), ("seed", u64), ("ev_hash", u64),
        ("c_key", (u32 * 4) * NP), ("river_tedashi", u32 * NP),
        ("score", i32 * NP),
        ("riichi_sticks", u32), ("turn_count", u32),
        ("step_count", u32), ("ev_count", u32), ("ev_words", u32),
        ("hand_len", u8 * NP), ("meld_tiles", ((u8 * 4) * 4) * NP),
        ("meld_type", (u8 * 4) * NP),
        ("n_melds", u8 * NP), ("n_river", u8 * NP), ("flags", u8 * NP),
        ("forbidden", (u8 * 2) * NP),
        ("wall_len", u8), ("wall_top", u8), ("rinshan_draw_count", u8),
        ("pending_kan_dora_count", u8), ("drawable_count", u8), ("n_dora", u8), ("dora_ind", u8 * 5), ("phase", u8),
        ("current_player", u8), ("oya", u8), ("honba", u8), ("kyoku_idx", u8),
        ("round_wind", u8), ("is_done", u8), ("needs_tsumo", u8), ("is_first_turn", u8),
        ("is_rinshan_flag", u8), ("riichi_pending_acceptance", u8), ("drawn_tile", u8), ("last_discard_pid", u8),
        ("last_discard_tile", u8), ("pending_kan_pid", u8), ("pending_kan_type", u8), ("pending_kan_tile", u8),
        ("active_mask", u8), ("last_error", u8), ("game_mode", u8), ("rule_bits", u8),
        ("overflow", u8), ("pending_init", u8 * 3), ("n_claims", u8 * NP),
        ("pending_tail", u8 * 2), ("is_after_kan", u8), ("hot_reserved", u8 * 9),
        # cold part
        ("wall", u8 * 136), ("river", (u8 * RIVER_CAP) * NP), ("claims", (u32 * MAX_CLAIMS) * NP),
        ("hand_index", u64), ("river_riichi", u32 * NP), ("score_delta", i32 * NP), ("meld_from", (u8 * 4) * NP), ("meld_called", (u8 * 4) * NP),
        ("pao", (u8 * 2) * NP), ("kyoku_count", u32), ("riichi_decl_idx", u8 * NP), ("riichi_sutehai", u8 * NP),
        ("last_tedashi", u8 * NP), ("n_kita", u8 * NP), ("reserved", u8 * 4),
    ]


HOT_BYTES = 544
assert GameState.wall.offset == HOT_BYTES and C.sizeof(GameState) % 16 == 0


class HandQuery(C.Structure):
    _fields_ = [
        ("tiles", u8 * 14), ("n_tiles", u8), ("n_melds", u8), ("meld_type", u8 * 4), ("meld_tiles", (u8 * 4) * 4),
        ("win_tile", u8), ("n_dora", u8), ("n_ura", u8), ("dora_ind", u8 * 5), ("ura_ind", u8 * 5),
        ("player_wind", u8), ("round_wind", u8), ("honba", u8), ("cond", u16), ("sanma", u8), ("kita_count", u8),
    ]


class HandResult(C.Structure):
    _fields_ = [
        ("yaku_mask", u64), ("wait_mask", u64), ("ron_agari", u32), ("tsumo_agari_oya", u32), ("tsumo_agari_ko", u32),
        ("is_win", u8), ("yakuman", u8), ("has_win_shape", u8), ("han", u8), ("fu", u8), ("shanten", i8),
        ("shanten13", i8), ("n_yaku", u8), ("_pad", u8 * 4),
    ]


class MjaiEvent(C.Structure):
    _fields_ = [("type", u8), ("actor", u8), ("target", u8), ("pai", u8), ("n_consumed", u8), ("consumed", u8 * 4), ("bakaze", u8),
                ("kyoku", u8), ("honba", u8), ("oya", u8), ("dora_marker", u8), ("tehai_len", u8 * NP), ("tehais", (u8 * 14) * NP),
                ("_pad", u8 * 3), ("kyotaku", u32), ("scores", i32 * NP)]


# replay ingestion (include/riichienv_b200.h: rv_log_action_type, rv_hule, rv_log_action, rv_log_kyoku)
LA_NONE, LA_DISCARD, LA_DEAL, LA_CHI_PENG_GANG, LA_ANGANG_ADDGANG, LA_DORA, LA_HULE, LA_NOTILE, LA_BABEI, LA_LIUJU = range(10)
LOG_MAX_DORAS = 8


class Hule(C.Structure):
    _fields_ = [("seat", u8), ("hu_tile", u8), ("zimo", u8), ("yiman", u8), ("n_li_doras", u8), ("li_doras", u8 * 5), ("_pad", u8 * 2),
                ("count", u32), ("fu", u32), ("point_rong", u32), ("point_zimo_qin", u32), ("point_zimo_xian", u32), ("fans", u64)]


class LogAction(C.Structure):
    _fields_ = [("type", u8), ("seat", u8), ("tile", u8), ("flags", u8), ("meld_type", u8), ("n_tiles", u8), ("tiles", u8 * 4),
                ("froms", u8 * 4), ("n_hule", u8), ("_pad", u8), ("hules", Hule * 3)]


class LogKyoku(C.Structure):
    _fields_ = [("np", u8), ("chang", u8), ("ju", u8), ("ben", u8), ("liqibang", u8), ("left_tile_count", u8), ("n_doras", u8),
                ("n_ura_doras", u8), ("doras", u8 * LOG_MAX_DORAS), ("ura_doras", u8 * LOG_MAX_DORAS), ("hand_len", u8 * NP),
                ("hands", (u8 * 14) * NP), ("wliqi", u8 * NP), ("oya", u8), ("oya_drawn_tile", u8), ("has_game_end_scores", u8),
                ("_pad", u8 * 3), ("rule_bits", u32), ("n_actions", i32), ("scores", i32 * NP), ("end_scores", i32 * NP),
                ("game_end_scores", i32 * NP)]


class LogActionAux(C.Structure):
    _fields_ = [("n_doras", u8), ("doras", u8 * LOG_MAX_DORAS), ("left_tile_count", u8), ("tile_raw_id", u8), ("_pad", u8)]


class WinContext(C.Structure):
    _fields_ = [("query", HandQuery), ("expected_yaku", C.c_uint64), ("expected_han", C.c_uint32), ("expected_fu", C.c_uint32),
                ("round", C.c_int32), ("action", C.c_int32), ("seat", u8), ("meld_from", i8 * 4), ("meld_called", u8 * 4), ("_pad", u8 * 7)]


class RunStats(C.Structure):
    _fields_ = [("games", C.c_int64), ("games_done", C.c_int64), ("env_steps", C.c_int64), ("rounds", C.c_int64),
                ("score_sum", C.c_int64 * NP), ("rank_hist", (C.c_int64 * NP) * NP)]


def state_fields_equal(a: "GameState", b: "GameState", skip=()):
    """Field-by-field comparison; returns list of differing field names."""
    diff = []
    for name, _ in GameState._fields_:
        if name in skip or name in ("reserved", "hot_reserved"):
            continue
        va, vb = getattr(a, name), getattr(b, name)
        if name == "claims":  # entries past n_claims are dead storage
            for p in range(NP):
                n = min(a.n_claims[p], MAX_CLAIMS)
                if list(va[p][:n]) != list(vb[p][:n]):
                    diff.append(name)
                    break
            continue
        if isinstance(va, C.Array):
            if bytes(va) != bytes(vb):
                diff.append(name)
        elif va != vb:
            diff.append(name)
    return diff
