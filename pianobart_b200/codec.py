"""Octuple <-> MIDI codec and the (N, 1024, 8) dataset format either side of the hot path.

Host-side integer work on variable-length note lists, mirroring reference `Data/data_generation/convert.py`:

    score_to_octuple()      MIDI_to_encoding          convert.py:157-252   notes -> sorted 8-tuples (+ task label)
    octuple_to_score()      encoding_to_MIDI          convert.py:256-318   8-tuples -> notes, time signatures, tempi
    pad_segment()           padding                   convert.py:321-333
    split_by_bar_limit()    the max_bar windows of F  convert.py:420-445
    pack_rows()             data_split                convert.py:575-580   flat rows -> (m, 1024, width) blocks
    segments_for_task()     F's per-task outputs      convert.py:446-508
    read_midi / write_midi  Standard MIDI File I/O so that the codec works without miditoolkit (absent from this image)

The reference operates on miditoolkit objects; everything here is duck-typed on the same attribute names
(`ticks_per_beat`, `instruments[i].notes[j].start/end/pitch/velocity`, `.program`, `.is_drum`, `.name`,
`time_signature_changes[i].time/numerator/denominator`, `tempo_changes[i].time/tempo`), so a miditoolkit `MidiFile` can be
passed in unchanged and `Score` (below) can stand in for it.  Reference quirks are kept because they define the data format the
model was trained on: drums are written as program 129 / pitch + 256 by the encoder although the decoder tests for program
128; Python's round-half-even maps ticks to positions; the most frequent time signature of a bar wins with CPython's set
iteration order breaking ties.  Parity: tests/test_codec.py against tests/golden/codec.npz, recorded by executing the
reference functions on synthetic scores (tools/make_golden.py::golden_codec, miditoolkit stubbed by containers).
"""
import math
import struct
from dataclasses import dataclass, field
from typing import List

import numpy as np

POS_RESOLUTION = 16            # positions per beat (quarter note)
MAX_BAR = 255
VELOCITY_QUANT = 4
TEMPO_QUANT = 12               # 2 ** (1 / 12) steps
MIN_TEMPO, MAX_TEMPO = 16, 256
DURATION_MAX = 8               # 2 ** 8 beats
MAX_TS_DENOMINATOR = 6         # x/1 ... x/64
MAX_NOTES_PER_BAR = 2
BEAT_NOTE_FACTOR = 4
TRUNC_POS = 2 ** 16
MAX_WINDOW = 1024
TOKENS_PER_NOTE = 8
# largest real value of every attribute (Bar, Pos, Program, Pitch, Duration, Velocity, TimeSig, Tempo); the special rows are
# boundary + 1 (<PAD>), + 2 (<MASK>), + 3 (<SOS>), + 4 (<EOS>)
TOKEN_BOUNDARY = (255, 127, 128, 255, 127, 31, 253, 48)
PAD_ROW = tuple(b + 1 for b in TOKEN_BOUNDARY)
EOS_ROW = tuple(b + 4 for b in TOKEN_BOUNDARY)
MELODY_MAP = {'MELODY': 0, 'BRIDGE': 1, 'PIANO': 2, 'OTHER': 3}
VELOCITY_MAP = {'pp': 0, 'p': 1, 'mp': 2, 'mf': 3, 'f': 4, 'ff': 5, 'OTHER': 6}
EMOTION_MAP = {'HVHA': 0, 'HVLA': 1, 'LVHA': 2, 'LVLA': 3}


# ------------------------------------------------------------------ containers (stand-ins for miditoolkit's)
@dataclass
class Note:
    start: int
    end: int
    pitch: int
    velocity: int


@dataclass
class Instrument:
    program: int = 0
    is_drum: bool = False
    name: str = ''
    notes: List[Note] = field(default_factory=list)


@dataclass
class TimeSignature:
    numerator: int
    denominator: int
    time: int


@dataclass
class TempoChange:
    tempo: float
    time: int


@dataclass
class Score:
    ticks_per_beat: int = 480
    instruments: List[Instrument] = field(default_factory=list)
    time_signature_changes: List[TimeSignature] = field(default_factory=list)
    tempo_changes: List[TempoChange] = field(default_factory=list)


# ------------------------------------------------------------------ attribute codes (convert.py:78-137)
def _build_tables():
    ts_index, ts_values = {}, []
    for i in range(MAX_TS_DENOMINATOR + 1):
        for j in range(1, (2 ** i) * MAX_NOTES_PER_BAR + 1):
            ts_index[(j, 2 ** i)] = len(ts_values)
            ts_values.append((j, 2 ** i))
    # durations: 16 linear steps per octave of length, step size doubling every octave
    enc, dec = [], []
    for octave in range(DURATION_MAX):
        for _ in range(POS_RESOLUTION):
            dec.append(len(enc))
            enc.extend([len(dec) - 1] * (2 ** octave))
    return ts_index, ts_values, enc, dec


_TS_INDEX, _TS_VALUES, _DUR_ENC, _DUR_DEC = _build_tables()


def timesig_to_code(ts):
    if ts not in _TS_INDEX:
        raise ValueError('unsupported time signature: ' + str(ts))
    return _TS_INDEX[ts]


def code_to_timesig(code):
    return _TS_VALUES[code]


def duration_to_code(d):
    return _DUR_ENC[d] if d < len(_DUR_ENC) else _DUR_ENC[-1]


def code_to_duration(code):
    return _DUR_DEC[code] if code < len(_DUR_DEC) else _DUR_DEC[-1]


def velocity_to_code(v):
    return v // VELOCITY_QUANT


def code_to_velocity(code):
    return code * VELOCITY_QUANT + VELOCITY_QUANT // 2


def tempo_to_code(bpm):
    bpm = min(max(bpm, MIN_TEMPO), MAX_TEMPO)
    return round(math.log2(bpm / MIN_TEMPO) * TEMPO_QUANT)


def code_to_tempo(code):
    return 2 ** (code / TEMPO_QUANT) * MIN_TEMPO


def reduce_time_signature(numerator, denominator):
    """convert.py:131-142: halve over-fine denominators, then split bars longer than MAX_NOTES_PER_BAR whole notes."""
    while denominator > 2 ** MAX_TS_DENOMINATOR and denominator % 2 == 0 and numerator % 2 == 0:
        denominator //= 2
        numerator //= 2
    while numerator > MAX_NOTES_PER_BAR * denominator:
        for f in range(2, numerator + 1):
            if numerator % f == 0:
                numerator //= f
                break
    return numerator, denominator


def _measure_length(ts_code):
    num, den = code_to_timesig(ts_code)
    return num * BEAT_NOTE_FACTOR * POS_RESOLUTION // den


# ------------------------------------------------------------------ score -> Octuple rows
def score_to_octuple(score, task='pretrain'):
    """convert.py:157-252.  Returns the sorted list of (bar, pos, program, pitch, duration, velocity, timesig, tempo) tuples of
    every note (9-tuples with the label for task 'melody' / 'velocity'); [] for a score without notes."""
    tpb = score.ticks_per_beat

    def pos_of(t):
        return round(t * POS_RESOLUTION / tpb)       # (Python's round-half-even, like the reference)

    starts = [pos_of(n.start) for inst in score.instruments for n in inst.notes]
    if not starts:
        return []
    n_pos = min(max(starts) + 1, TRUNC_POS)
    ts_code = np.full(n_pos, timesig_to_code(reduce_time_signature(4, 4)), dtype=np.int64)    # MIDI default 4/4
    tp_code = np.full(n_pos, tempo_to_code(120.0), dtype=np.int64)                          # MIDI default 120 BPM

    def paint(arr, changes, code_of):
        for i, ch in enumerate(changes):
            lo = pos_of(ch.time)
            hi = pos_of(changes[i + 1].time) if i + 1 < len(changes) else n_pos
            hi = min(hi, n_pos)
            if lo < hi:                              # (the code is only evaluated for positions inside the piece, as in the
                arr[lo:hi] = code_of(ch)             #  reference: an unsupported signature after the last note is not an error)

    paint(ts_code, score.time_signature_changes, lambda c: timesig_to_code(reduce_time_signature(c.numerator, c.denominator)))
    paint(tp_code, score.tempo_changes, lambda c: tempo_to_code(c.tempo))
    bar_of = np.empty(n_pos, dtype=np.int64)
    pos_in_bar = np.empty(n_pos, dtype=np.int64)
    bar, cnt, length = 0, 0, None
    for j in range(n_pos):
        if cnt == 0:
            length = _measure_length(int(ts_code[j]))
        bar_of[j], pos_in_bar[j] = bar, cnt
        cnt += 1
        if cnt >= length:
            if cnt != length:
                raise ValueError('invalid time signature change: pos = {}'.format(j))
            cnt -= length
            bar += 1
    rows = []
    for inst in score.instruments:
        program = 129 if inst.is_drum else inst.program          # (max_inst + 1: see the module docstring)
        for n in inst.notes:
            p = pos_of(n.start)
            if p >= TRUNC_POS:
                continue
            row = (int(bar_of[p]), int(pos_in_bar[p]), program, n.pitch + 256 if inst.is_drum else n.pitch,
                   duration_to_code(pos_of(n.end) - p), velocity_to_code(n.velocity), int(ts_code[p]), int(tp_code[p]))
            if task == 'melody':
                row += (MELODY_MAP.get(inst.name, MELODY_MAP['OTHER']),)
            elif task == 'velocity':
                if 0 <= n.velocity <= 15:
                    label = 0
                elif 112 <= n.velocity <= 127:
                    label = 5
                else:
                    label = (n.velocity - 32) // 16 + 1
                if not 0 <= label <= 5:
                    raise ValueError('velocity label out of range')
                row += (label,)
            rows.append(row)
    rows.sort()
    return rows


# ------------------------------------------------------------------ Octuple rows -> score
def octuple_to_score(rows, ticks_per_beat=480):
    """convert.py:256-318.  rows: iterable of 8-tuples (special rows must have been cut off, e.g. by
    postprocess.octuple_truncate).  Returns a Score with one Instrument per program that has notes."""
    rows = [tuple(int(x) for x in r) for r in rows]
    n_bars = max(r[0] for r in rows) + 1
    per_bar = [[] for _ in range(n_bars)]
    for r in rows:
        per_bar[r[0]].append(r[6])
    bar_ts = [max(set(v), key=v.count) if v else None for v in per_bar]      # most frequent; ties: CPython set order
    for i in range(n_bars):
        if bar_ts[i] is None:
            bar_ts[i] = timesig_to_code(reduce_time_signature(4, 4)) if i == 0 else bar_ts[i - 1]
    bar_start, cur = [0] * n_bars, 0
    for i in range(n_bars):
        bar_start[i] = cur
        if 0 <= bar_ts[i] < len(_TS_VALUES):          # (a special-token code adds no length, as the reference's try/except)
            cur += _measure_length(bar_ts[i])
    n_tempo_pos = cur + max(r[1] for r in rows)
    tempo_votes = [[] for _ in range(n_tempo_pos)]
    for r in rows:
        k = bar_start[r[0]] + r[1]
        if -n_tempo_pos <= k < n_tempo_pos:
            tempo_votes[k].append(r[7])
    tempo_at = [round(sum(v) / len(v)) if v else None for v in tempo_votes]
    for i in range(n_tempo_pos):
        if tempo_at[i] is None:
            tempo_at[i] = tempo_to_code(120.0) if i == 0 else tempo_at[i - 1]

    def tick(bar, pos):
        return (bar_start[bar] + pos) * ticks_per_beat // POS_RESOLUTION

    insts = [Instrument(program=0 if i == 128 else i, is_drum=(i == 128), name=str(i)) for i in range(129)]
    for r in rows:
        start = tick(r[0], r[1])
        program = r[2]
        pitch = r[3] - 128 if program == 128 else r[3]
        dur = tick(0, code_to_duration(r[4])) or 1
        if -129 <= program < 129:                     # (list index semantics of the reference; other programs are dropped)
            insts[program].notes.append(Note(start=start, end=start + dur, pitch=pitch, velocity=code_to_velocity(r[5])))
    out = Score(ticks_per_beat=ticks_per_beat, instruments=[i for i in insts if i.notes])
    last = None
    for i in range(n_bars):
        if bar_ts[i] != last:
            if not 0 <= bar_ts[i] < len(_TS_VALUES):
                continue
            num, den = code_to_timesig(bar_ts[i])
            out.time_signature_changes.append(TimeSignature(numerator=num, denominator=den, time=tick(i, 0)))
            last = bar_ts[i]
    last = None
    for i in range(n_tempo_pos):
        if tempo_at[i] != last:
            out.tempo_changes.append(TempoChange(tempo=code_to_tempo(tempo_at[i]), time=tick(0, i)))
            last = tempo_at[i]
    return out


# ------------------------------------------------------------------ dataset blocks
def pad_segment(rows, window=MAX_WINDOW, last=False):
    """convert.py:321-333: <PAD> rows up to `window`; a longer segment keeps its first (or last) window - 1 rows + <EOS>."""
    rows = list(rows)
    if len(rows) > window:
        rows = rows[1 - window:] if last else rows[:window - 1]
        rows.append(EOS_ROW)
        return rows
    rows.extend([PAD_ROW] * (window - len(rows)))
    return rows


def split_by_bar_limit(rows):
    """convert.py:420-445: a piece longer than MAX_BAR bars is cut where the bar index passes k * MAX_BAR; bar indices of the
    later parts restart (minus (k - 1) * MAX_BAR + 1); every part ends with <EOS>."""
    parts, start, k = [], 0, 1

    def rebased(seg, k):
        if k > 1:
            off = MAX_BAR * (k - 1) + 1
            seg = [(r[0] - off,) + tuple(r[1:]) for r in seg]
        return list(seg) + [EOS_ROW]

    for i, r in enumerate(rows):
        if r[0] > MAX_BAR * k:
            parts.append(rebased(rows[start:i], k))
            start, k = i, k + 1
    parts.append(rebased(rows[start:], k))
    return parts


def pack_rows(data, fill=PAD_ROW, width=TOKENS_PER_NOTE):
    """convert.py:575-580: flat rows -> (m, 1024, width), the tail block filled with `fill` (always at least one fill row)."""
    data = np.asarray(data)
    if data.ndim == 1:
        data = data.reshape(-1, width) if width > 1 else data
    m = data.shape[0] // MAX_WINDOW + 1
    n_fill = m * MAX_WINDOW - data.shape[0]
    fill_rows = np.asarray([fill] * n_fill).reshape((n_fill,) + data.shape[1:]) if n_fill else data[:0]
    return np.concatenate([data, fill_rows.astype(data.dtype, copy=False)], axis=0).reshape(m, MAX_WINDOW, width)


def segments_for_task(rows, task='pretrain', pad=True, label=None):
    """convert.py:446-508 for one piece (after score_to_octuple): list of outputs per bar-limited part -
    'pretrain': padded (or raw) segments; 'composer' / 'emotion': (padded segment, label); 'melody' / 'velocity':
    (8-tuples, per-note labels with OTHER for the <EOS> row); 'generate': (prompt, continuation) pairs cut at a bar boundary."""
    out = []
    for seg in split_by_bar_limit(rows):
        if task == 'generate':
            half = MAX_WINDOW - 1 if len(seg) >= 2 * MAX_WINDOW else len(seg) // 2 - 1
            head = seg[:half]
            if not head:
                raise ValueError('piece too short to split into prompt and continuation')
            cut = 0
            for cut, r in enumerate(head):
                if r[0] >= head[-1][0]:
                    break
            prompt, cont = list(seg[:cut]), list(seg[cut:])
            prompt.append(EOS_ROW)
            prompt, cont = pad_segment(prompt), pad_segment(cont)
            if sum(1 for r in prompt if r[0] == EOS_ROW[0]) != 1:
                continue
            out.append((prompt, cont))
        elif task == 'pretrain':
            out.append(pad_segment(seg) if pad else seg)
        elif task in ('composer', 'emotion'):
            out.append((pad_segment(seg), label))
        elif task in ('melody', 'velocity'):
            other = MELODY_MAP['OTHER'] if task == 'melody' else VELOCITY_MAP['OTHER']
            labels = [r[-1] if len(r) == 9 else other for r in seg]
            out.append(([tuple(r[:TOKENS_PER_NOTE]) for r in seg], labels))
        else:
            raise ValueError('unknown task ' + str(task))
    return out


# ------------------------------------------------------------------ Standard MIDI File I/O (format 0 / 1)
def _read_varlen(buf, i):
    v = 0
    while True:
        b = buf[i]
        i += 1
        v = (v << 7) | (b & 0x7f)
        if not b & 0x80:
            return v, i


def read_midi(path):
    """Minimal SMF reader (format 0 / 1): what `miditoolkit.midi.parser.MidiFile(path)` gives convert.py:336 / demo.py:62 -
    ticks_per_beat, instruments with notes in ticks, set-tempo and time-signature changes.  miditoolkit is absent from this
    image (and unpinned by the reference), so its published loading rule - the one it shares with pretty_midi - is restated:
    a note-off (or note-on with velocity 0) closes EVERY open note-on of its (channel, pitch) that started on an earlier
    tick, note-ons of the same tick stay open if something was closed (else they are dropped with the key); a note belongs
    to the channel's program at its note-OFF; instruments are keyed (program, channel, track) in order of their first closed
    note, channel 10 is drums, the track name is the instrument name.  Running status, sysex and unknown meta events are
    skipped over.  (Parity with the library itself is unpinned: no copy of it exists here.)"""
    with open(path, 'rb') as f:
        buf = f.read()
    if buf[:4] != b'MThd':
        raise ValueError('not a Standard MIDI File')
    hlen, fmt, ntrk, div = struct.unpack('>IHHH', buf[4:14])
    if div & 0x8000:
        raise ValueError('SMPTE time division is not supported')
    score = Score(ticks_per_beat=div)
    i = 8 + hlen
    for trk in range(ntrk):
        if buf[i:i + 4] != b'MTrk':
            raise ValueError('bad track chunk')
        tlen = struct.unpack('>I', buf[i + 4:i + 8])[0]
        j, end = i + 8, i + 8 + tlen
        i = end
        t, status = 0, 0
        program = [0] * 16
        open_notes = {}
        insts = {}           # (program, channel) -> Instrument, insertion-ordered (one track at a time)
        name = ''
        while j < end:
            dt, j = _read_varlen(buf, j)
            t += dt
            b = buf[j]
            if b & 0x80:
                status = b
                j += 1
            if status == 0xff:
                kind = buf[j]
                ln, j = _read_varlen(buf, j + 1)
                data = buf[j:j + ln]
                j += ln
                if kind == 0x51 and ln == 3:
                    score.tempo_changes.append(TempoChange(tempo=60e6 / int.from_bytes(data, 'big'), time=t))
                elif kind == 0x58 and ln >= 2:
                    score.time_signature_changes.append(TimeSignature(numerator=data[0], denominator=2 ** data[1], time=t))
                elif kind == 0x03:
                    name = data.decode('latin1')
                    for inst in insts.values():
                        inst.name = name
                continue
            if status in (0xf0, 0xf7):
                ln, j = _read_varlen(buf, j)
                j += ln
                continue
            hi, ch = status & 0xf0, status & 0x0f
            if hi in (0xc0, 0xd0):
                if hi == 0xc0:
                    program[ch] = buf[j]
                j += 1
                continue
            d1, d2 = buf[j], buf[j + 1]
            j += 2
            if hi == 0x90 and d2 > 0:
                open_notes.setdefault((ch, d1), []).append((t, d2))
            elif hi == 0x80 or (hi == 0x90 and d2 == 0):
                q = open_notes.get((ch, d1))
                if q is not None:
                    close = [(t0, vel) for t0, vel in q if t0 != t]
                    keep = [(t0, vel) for t0, vel in q if t0 == t]
                    for t0, vel in close:
                        key = (program[ch], ch)
                        if key not in insts:
                            insts[key] = Instrument(program=program[ch], is_drum=(ch == 9), name=name)
                        insts[key].notes.append(Note(start=t0, end=t, pitch=d1, velocity=vel))
                    if close and keep:
                        open_notes[(ch, d1)] = keep
                    else:
                        del open_notes[(ch, d1)]
        score.instruments.extend(insts.values())
    score.tempo_changes.sort(key=lambda c: c.time)
    score.time_signature_changes.sort(key=lambda c: c.time)
    return score


def _varlen(v):
    out = [v & 0x7f]
    v >>= 7
    while v:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def write_midi(score, path):
    """Format-1 SMF: track 0 carries tempo / time-signature changes, one track per instrument (drums on channel 10)."""
    def chunk(events):
        events.sort(key=lambda e: (e[0], e[1]))
        body, t = bytearray(), 0
        for when, _, data in events:
            body += _varlen(when - t) + data
            t = when
        body += b'\x00\xff\x2f\x00'
        return b'MTrk' + struct.pack('>I', len(body)) + bytes(body)

    meta = []
    for c in score.tempo_changes:
        meta.append((int(c.time), 0, b'\xff\x51\x03' + int(round(60e6 / c.tempo)).to_bytes(3, 'big')))
    for c in score.time_signature_changes:
        meta.append((int(c.time), 0, b'\xff\x58\x04' + bytes([c.numerator, int(math.log2(c.denominator)), 24, 8])))
    tracks = [chunk(meta)]
    free = [c for c in range(16) if c != 9]
    for k, inst in enumerate(score.instruments):
        ch = 9 if inst.is_drum else free[k % len(free)]
        ev = [(0, 0, bytes([0xc0 | ch, inst.program & 0x7f]))]
        for n in inst.notes:
            ev.append((int(n.start), 2, bytes([0x90 | ch, n.pitch & 0x7f, max(1, min(127, n.velocity))])))
            ev.append((int(n.end), 1, bytes([0x80 | ch, n.pitch & 0x7f, 0])))
        tracks.append(chunk(ev))
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 1, len(tracks), score.ticks_per_beat) + b''.join(tracks))
