"""Token -> notes -> MIDI (inpaintnet_b200/score.py): the reference's tensor_to_score
(DatasetManager/the_session/folk_dataset.py:472-502) restated without music21, checked on worked examples of its rules
and by parsing the written Standard MIDI File back."""
from fractions import Fraction as F

import torch

from inpaintnet_b200 import score as S
from inpaintnet_b200.data import SyntheticFolkDataset


def _parse_midi(data):
    assert data[:4] == b"MThd" and int.from_bytes(data[4:8], "big") == 6
    fmt, ntrk, ppq = (int.from_bytes(data[8 + 2 * i:10 + 2 * i], "big") for i in range(3))
    assert data[14:18] == b"MTrk"
    n = int.from_bytes(data[18:22], "big")
    body, i, t, events = data[22:22 + n], 0, 0, []
    assert 22 + n == len(data)
    while i < len(body):
        delta = 0
        while True:
            b = body[i]; i += 1
            delta = (delta << 7) | (b & 0x7F)
            if not b & 0x80:
                break
        t += delta
        st = body[i]
        if st == 0xFF:
            ln = body[i + 2]
            events.append((t, "meta", body[i + 1], bytes(body[i + 3:i + 3 + ln])))
            i += 3 + ln
        else:
            events.append((t, "on" if st & 0xF0 == 0x90 else "off", body[i + 1], body[i + 2]))
            i += 3
    return fmt, ntrk, ppq, events


def test_tick_grid_and_names():
    assert S.tick_durations() == [F(1, 4), F(1, 12), F(1, 6), F(1, 6), F(1, 12), F(1, 4)] and sum(S.tick_durations()) == 1
    assert S.name_to_midi("C4") == 60 and S.name_to_midi("F#5") == 78 and S.name_to_midi("B-3") == 58
    assert all(S.name_to_midi(s) is None for s in ("rest", "__", "START", "END", "OOR", "XX"))
    assert all(S.name_to_midi(S.midi_to_name(m)) == m for m in range(40, 100))
    v = S.default_vocabulary(64)
    assert len(v) == 64 == len(set(v)) and v[:4] == ["rest", "__", "START", "END"] and v[4] == "G3"


def test_slur_extends_and_special_symbols_are_rests():
    ds = SyntheticFolkDataset(num_notes=64)
    n2i = ds.note2index_dicts[0]
    c4, d4, slur, rest, start = n2i["C4"], n2i["D4"], n2i["__"], n2i["rest"], n2i["START"]
    # one beat: C4 held over three ticks (1/4 + 1/12 + 1/6), D4 over two (1/6 + 1/12), a rest on the last sixteenth
    sc = ds.tensor_to_score(torch.tensor([[c4, slur, slur, d4, slur, rest]]))
    assert [(n.name, n.quarter_length) for n in sc.notes] == [("C4", F(1, 2)), ("D4", F(1, 4)), ("rest", F(1, 4))]
    # leading slurs lengthen the initial (zero-length) rest; a special symbol sounds as a rest; repeated pitches re-articulate
    sc = ds.tensor_to_score(torch.tensor([slur, slur, c4, c4, start, slur]))
    assert [(n.is_rest, n.quarter_length) for n in sc.notes] == [(True, F(1, 3)), (False, F(1, 6)), (False, F(1, 6)), (True, F(1, 3))]
    assert sc.quarter_length == 1


def test_total_length_and_midi_round_trip(tmp_path):
    ds = SyntheticFolkDataset(num_notes=64)
    g = torch.Generator().manual_seed(3)
    tokens = torch.randint(0, 64, (1, 16 * 24), generator=g)
    tokens[0, ::3] = ds.note2index_dicts[0]["__"]
    sc = ds.tensor_to_score(tokens)
    assert sc.quarter_length == 16 * 4                      # 16 bars of 4 beats, one quarter length per beat
    path = sc.write("midi", fp=str(tmp_path / "lead.mid"))
    fmt, ntrk, ppq, ev = _parse_midi(open(path, "rb").read())
    assert (fmt, ntrk, ppq) == (0, 1, 480)
    assert ev[0][1:3] == ("meta", 0x51) and int.from_bytes(ev[0][3], "big") == 500000
    assert ev[-1][1:3] == ("meta", 0x2F) and ev[-1][0] == 16 * 4 * 480
    ons = [e for e in ev if e[1] == "on"]
    offs = [e for e in ev if e[1] == "off"]
    sounding = [n for n in sc.notes if not n.is_rest]
    assert len(ons) == len(offs) == len(sounding)
    t = 0
    k = 0
    for n in sc.notes:                                      # every note starts and ends on its own pulse
        if not n.is_rest:
            assert ons[k][0] == t and ons[k][2] == n.midi and offs[k][0] == t + n.quarter_length * 480
            k += 1
        t += int(n.quarter_length * 480)
