"""TEST INFRASTRUCTURE (oracle): a plain-Python restatement of WinResultContextIterator::do_next (riichienv-core/src/replay/mod.rs:1741-2093)
over the shim's action views — the checker for rv_replay_win_contexts (csrc/replay.cpp).  Returns, per Hule, the tuple
(seat, tiles, melds, win tile, dora markers, ura markers, condition bits, player wind, round wind, kita count)."""


class A:  # rv_log_action_type (include/riichienv_b200.h), restated so that the checker imports nothing of the product
    LA_DISCARD, LA_DEAL, LA_CHI_PENG_GANG, LA_ANGANG_ADDGANG, LA_DORA, LA_HULE, LA_NOTILE, LA_BABEI, LA_LIUJU = range(1, 10)


PON, DAIMINKAN, ANKAN, KAKAN = 1, 2, 3, 4


def _remove(hand, t):  # TileConverter::match_and_remove_u8 (2243-2255)
    if t in hand:
        hand.remove(t)
        return
    for x in hand:
        if x // 4 == t // 4:
            hand.remove(x)
            return


def walk(kyoku, wall=()):
    """kyoku: riichienv_b200.replay.LogKyoku; wall: LogKyoku.paishan as tile ids (empty = none)"""
    np_ = len(kyoku.scores)
    hands = [list(h) for h in kyoku.hands]
    melds = [[] for _ in range(4)]
    liqi, wliqi, ippatsu, rinshan, first = ([False] * 4 for _ in range(5))
    first = [True] * 4
    before_babei = [False] * 4
    last_kakan = last_babei = False
    kakan_tile = None
    doras = list(kyoku.doras)
    left = kyoku.left_tile_count
    dora_count, pending = 1, 0
    kita = [0] * 4
    wall = list(wall)
    out = []

    def sync():
        nonlocal dora_count, pending, doras
        if not wall:
            return
        if len(doras) > dora_count:
            dora_count, pending = len(doras), 0
        elif dora_count > len(doras):
            doras = [wall[len(wall) - 5 - 2 * i] for i in range(dora_count) if len(wall) >= 5 + 2 * i]

    def flush():
        nonlocal dora_count, pending
        if pending > 0:
            dora_count, pending = dora_count + pending, 0

    def after_kakan():
        nonlocal last_kakan, last_babei, kakan_tile
        if last_kakan:
            ippatsu[:] = [False] * 4
            first[:] = [False] * 4
            last_kakan = last_babei = False
            kakan_tile = None

    for a in kyoku._views:
        if a.type != A.LA_HULE:
            rinshan[:] = [False] * 4
            if a.type != A.LA_BABEI:
                last_babei = False
        s = a.seat
        if a.type == A.LA_DISCARD:
            after_kakan()
            if a.is_wliqi:
                wliqi[s] = ippatsu[s] = True
            if a.is_liqi:
                liqi[s] = ippatsu[s] = True
            else:
                ippatsu[s] = False
            first[s] = False
            _remove(hands[s], a.tile)
            if a.doras is not None:
                doras = list(a.doras)
            flush()
            sync()
        elif a.type == A.LA_DEAL:
            after_kakan()
            hands[s].append(a.tile)
            if a.left_tile_count is not None:
                left = a.left_tile_count
            elif left > 0:
                left -= 1
            if a.doras is not None:
                doras = list(a.doras)
                rinshan[s] = True
            sync()
        elif a.type == A.LA_CHI_PENG_GANG:
            rinshan[:] = [False] * 4
            ippatsu[:] = [False] * 4
            first[:] = [False] * 4
            last_kakan = last_babei = False
            kakan_tile = None
            for t, f in zip(a.tiles, a.froms):
                if f == s:
                    _remove(hands[s], t)
            melds[s].append([a.meld_type, list(a.tiles)])
            if a.meld_type == DAIMINKAN:
                rinshan[s] = True
                flush()
                pending += 1
        elif a.type == A.LA_DORA:
            if not wall:
                doras.append(a.tile)
            else:
                dora_count += 1
                pending = max(0, pending - 1)
                sync()
        elif a.type == A.LA_ANGANG_ADDGANG:
            rinshan[:] = [False] * 4
            flush()
            if a.meld_type == ANKAN:
                ippatsu[:] = [False] * 4
                first[:] = [False] * 4
                last_kakan = last_babei = False
                kakan_tile = None
                k34 = kyoku._aux[kyoku._views.index(a)].tile_raw_id if kyoku._aux is not None else 0
                for _ in range(4):
                    for x in hands[s]:
                        if x // 4 == k34:
                            hands[s].remove(x)
                            break
                melds[s].append([ANKAN, [4 * k34 + i for i in range(4)]])
                rinshan[s] = True
                if wall:
                    dora_count += 1
            else:
                last_kakan, kakan_tile = True, a.tiles[0]
                rinshan[s] = True
                for m in melds[s]:
                    if m[0] == PON and m[1][0] // 4 == a.tiles[0] // 4:
                        m[0] = KAKAN
                        m[1].append(a.tiles[0])
                        break
                else:
                    melds[s].append([a.meld_type, list(a.tiles)])
                _remove(hands[s], a.tiles[0])
                pending += 1
            sync()
        elif a.type == A.LA_BABEI:
            before_babei = list(ippatsu)
            ippatsu[:] = [False] * 4
            first[:] = [False] * 4
            last_babei = True
            for x in hands[s]:
                if x // 4 == 30:
                    hands[s].remove(x)
                    break
            kita[s] += 1
            rinshan[s] = True
        elif a.type == A.LA_HULE:
            for h in a.hules:
                w, zimo = h.seat, bool(h.zimo)
                chankan = (not zimo) and last_kakan and kakan_tile is not None and kakan_tile // 4 == h.hu_tile // 4
                ipp = before_babei[w] if (not zimo and last_babei) else ippatsu[w]
                tiles = list(hands[w]) + ([] if zimo else [h.hu_tile])
                if liqi[w]:
                    if h.n_li_doras != 0xFF:
                        ura = [h.li_doras[i] for i in range(h.n_li_doras)]
                    elif wall:
                        ura = [wall[len(wall) - 6 - 2 * i] for i in range(dora_count) if len(wall) >= 6 + 2 * i]
                    else:
                        ura = list(kyoku.ura_doras)
                else:
                    ura = []
                bits = (zimo * 1 | liqi[w] * 2 | wliqi[w] * 4 | ipp * 8 | (left == 0 and zimo and not rinshan[w]) * 16 |
                        (left == 0 and not zimo and not rinshan[w]) * 32 | rinshan[w] * 64 | chankan * 128 | (first[w] and zimo) * 256)
                out.append((w, tiles, [(m[0], list(m[1])) for m in melds[w]], h.hu_tile, list(doras)[:5], ura[:5], int(bits),
                            (w + np_ - kyoku.ju % np_) % np_, kyoku.chang & 3, kita[w]))
    return out
