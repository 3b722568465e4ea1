"""Token tensor -> notes -> Standard MIDI File, host-side post-processing of the
argmax token output.

The reference turns a lead-sheet token tensor into a music21 Score (DatasetManager/the_session/folk_dataset.py:472-502)
and lets music21 export it.  music21 is not available here, so the same conversion is restated on plain Python
objects: one token per tick, 6 ticks per beat on the reference's mixed sixteenth / triplet grid
(folk_data_helpers.py:22-29: tick values 0, 1/4, 1/3, 1/2, 2/3, 3/4 of a beat), the slur symbol `__` extends the
sounding note or rest, every other special symbol is a rest (DatasetManager/helpers.py:38-56).  `Score.write("midi",
fp)` writes a format-0 Standard MIDI File directly.
"""
from fractions import Fraction

SLUR_SYMBOL, START_SYMBOL, END_SYMBOL, OUT_OF_RANGE, PAD_SYMBOL, REST_SYMBOL = "__", "START", "END", "OOR", "XX", "rest"
TICK_VALUES = (Fraction(0), Fraction(1, 4), Fraction(1, 3), Fraction(1, 2), Fraction(2, 3), Fraction(3, 4))
_STEP = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}
_SHARP_NAMES = ("C", "C#", "D", "E-", "E", "F", "F#", "G", "G#", "A", "B-", "B")   # music21 spelling: '-' is a flat


def tick_durations(tick_values=TICK_VALUES):
    """Duration of every tick of a beat in quarter lengths (folk_dataset.py:72-79)."""
    d = [n - p for n, p in zip(tick_values[1:], tick_values[:-1])]
    return d + [1 - tick_values[-1]]


def midi_to_name(midi):
    return f"{_SHARP_NAMES[midi % 12]}{midi // 12 - 1}"


def name_to_midi(name):
    """'C4' -> 60, 'F#5' -> 78, 'B-3' -> 58; None for rests and the special symbols (helpers.py:38-56)."""
    if not isinstance(name, str) or name in (REST_SYMBOL, SLUR_SYMBOL, START_SYMBOL, END_SYMBOL, OUT_OF_RANGE, PAD_SYMBOL):
        return None
    step = name[0].upper()
    if step not in _STEP:
        return None
    i, alter = 1, 0
    while i < len(name) and name[i] in "#-b":
        alter += 1 if name[i] == "#" else -1
        i += 1
    try:
        octave = int(name[i:])
    except ValueError:
        return None
    return 12 * (octave + 1) + _STEP[step] + alter


def default_vocabulary(num_notes):
    """A folk-like symbol table for synthetic data: rest, slur, START, END, then the chromatic pitches from G3 (MIDI 55,
    the lower end of the reference's pitch range, folk_dataset.py:36) upwards."""
    names = [REST_SYMBOL, SLUR_SYMBOL, START_SYMBOL, END_SYMBOL]
    midi = 55
    while len(names) < num_notes:
        names.append(midi_to_name(midi))
        midi += 1
    return names[:num_notes]


class Note:
    """One sounding event: `midi` is None for a rest; `quarter_length` is an exact Fraction."""

    def __init__(self, name, quarter_length):
        self.name = name
        self.midi = name_to_midi(name)
        self.quarter_length = Fraction(quarter_length)

    @property
    def is_rest(self):
        return self.midi is None

    def __repr__(self):
        return f"Note({'rest' if self.is_rest else self.name!r}, {self.quarter_length})"


def _vlq(n):
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


class Score:
    """A monophonic lead: the flat list of notes / rests `tensor_to_score` produces."""

    PPQ = 480   # pulses per quarter: a multiple of 12, so every tick duration of the grid is a whole number of pulses

    def __init__(self, notes, tempo_bpm=120):
        self.notes = list(notes)
        self.tempo_bpm = tempo_bpm

    @property
    def quarter_length(self):
        return sum((n.quarter_length for n in self.notes), Fraction(0))

    def to_midi_bytes(self, velocity=80, channel=0):
        track = bytearray()
        us_per_quarter = int(round(60_000_000 / self.tempo_bpm))
        track += _vlq(0) + bytes([0xFF, 0x51, 0x03]) + us_per_quarter.to_bytes(3, "big")
        pending = 0   # pulses of silence accumulated since the last event
        for n in self.notes:
            pulses = n.quarter_length * self.PPQ
            assert pulses.denominator == 1, f"duration {n.quarter_length} is not on the {self.PPQ}-pulse grid"
            pulses = int(pulses)
            if n.is_rest:
                pending += pulses
                continue
            track += _vlq(pending) + bytes([0x90 | channel, n.midi & 0x7F, velocity])
            track += _vlq(pulses) + bytes([0x80 | channel, n.midi & 0x7F, 0])
            pending = 0
        track += _vlq(pending) + bytes([0xFF, 0x2F, 0x00])
        header = b"MThd" + (6).to_bytes(4, "big") + (0).to_bytes(2, "big") + (1).to_bytes(2, "big") + self.PPQ.to_bytes(2, "big")
        return header + b"MTrk" + len(track).to_bytes(4, "big") + bytes(track)

    def write(self, fmt="midi", fp=None):
        """music21-shaped call: score.write('midi', fp=path) -> path."""
        if fmt not in ("midi", "mid"):
            raise NotImplementedError(f"only MIDI export is provided (asked for {fmt!r}); music21 is not a dependency")
        if fp is None:
            raise ValueError("write('midi', fp=...) needs a file path")
        with open(fp, "wb") as f:
            f.write(self.to_midi_bytes())
        return fp

    def __len__(self):
        return len(self.notes)

    def __repr__(self):
        return f"Score({len(self.notes)} events, {self.quarter_length} quarter lengths)"


def tokens_to_score(tokens, index2note, slur_index, durations=None):
    """folk_dataset.py:472-502: a token that is not the slur starts a new note (or rest) and flushes the previous one;
    a slur lengthens what is sounding.  The sequence starts in a rest of length 0, so leading slurs become a rest."""
    durations = durations or tick_durations()
    sub = len(durations)
    notes = []
    dur = Fraction(0)
    name = REST_SYMBOL
    for tick, tok in enumerate(int(t) for t in tokens):
        if tok != slur_index:
            if dur > 0:
                notes.append(Note(name, dur))
            dur = durations[tick % sub]
            name = index2note[tok]
        else:
            dur += durations[tick % sub]
    notes.append(Note(name, dur))
    return Score(notes)
