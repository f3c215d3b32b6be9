"""Octuple <-> MIDI codec against fixtures recorded by executing the reference's convert.py
(tests/golden/codec.npz, tools/make_golden.py::golden_codec): MIDI_to_encoding for the pretrain / melody / velocity tasks,
encoding_to_MIDI, padding, data_split and the scalar code tables - bit-exact integer work.  CPU only."""
import numpy as np
import pytest

from pianobart_b200 import codec as K
from util import load_golden

NAMES = ['MELODY', 'BRIDGE', 'PIANO', 'x']


def _score(g, ci):
    sc = K.Score(ticks_per_beat=int(g['c%d_tpb' % ci]))
    for prog, drum, name in g['c%d_inst' % ci]:
        sc.instruments.append(K.Instrument(program=int(prog), is_drum=bool(drum), name=NAMES[name] if name >= 0 else 'drums'))
    for k, st, en, pitch, vel in g['c%d_notes' % ci]:
        sc.instruments[int(k)].notes.append(K.Note(int(st), int(en), int(pitch), int(vel)))
    sc.time_signature_changes = [K.TimeSignature(int(n), int(d), int(t)) for t, n, d in g['c%d_ts' % ci]]
    sc.tempo_changes = [K.TempoChange(float(v), int(t)) for t, v in zip(g['c%d_tempo_t' % ci], g['c%d_tempo_v' % ci])]
    return sc


def test_code_tables_match_reference():
    g = load_golden('codec')
    assert [K.duration_to_code(x) for x in range(0, 5000, 7)] == g['tab_d2e'].tolist()
    assert [K.code_to_duration(x) for x in range(0, 140)] == g['tab_e2d'].tolist()
    assert [K.tempo_to_code(x) for x in np.linspace(5, 400, 300)] == g['tab_b2e'].tolist()
    assert np.array_equal(np.array([K.code_to_tempo(x) for x in range(0, 49)]), g['tab_e2b'])
    got = [K.reduce_time_signature(n, d) for n in range(1, 40) for d in (1, 2, 4, 8, 16, 32, 64, 128)]
    assert np.array_equal(np.array(got), g['tab_tsr'])
    assert K.PAD_ROW == (256, 128, 129, 256, 128, 32, 254, 49) and K.EOS_ROW[0] == 259


@pytest.mark.parametrize('ci', range(5))
def test_score_to_octuple_matches_reference(ci):
    g = load_golden('codec')
    sc = _score(g, ci)
    for task in ('pretrain', 'melody', 'velocity'):
        got = np.array(K.score_to_octuple(sc, task), dtype=np.int64)
        assert np.array_equal(got, g['c%d_enc_%s' % (ci, task)]), task


@pytest.mark.parametrize('ci', range(5))
def test_octuple_to_score_matches_reference(ci):
    g = load_golden('codec')
    rows = g['c%d_enc_pretrain' % ci]
    back = K.octuple_to_score(rows)
    notes = np.array([[int(i.name), int(i.is_drum), i.program, n.start, n.end, n.pitch, n.velocity]
                      for i in back.instruments for n in i.notes], dtype=np.int64)
    assert np.array_equal(notes, g['c%d_dec_notes' % ci])
    ts = np.array([[c.time, c.numerator, c.denominator] for c in back.time_signature_changes], dtype=np.int64)
    assert np.array_equal(ts, g['c%d_dec_ts' % ci])
    assert [c.time for c in back.tempo_changes] == g['c%d_dec_tempo_t' % ci].tolist()
    assert np.array_equal(np.array([c.tempo for c in back.tempo_changes]), g['c%d_dec_tempo_v' % ci])


@pytest.mark.parametrize('ci', range(5))
def test_dataset_blocks_match_reference(ci):
    g = load_golden('codec')
    rows = [tuple(r) for r in g['c%d_enc_pretrain' % ci].tolist()]
    assert np.array_equal(np.array(K.pad_segment(rows[:1500])), g['c%d_pad' % ci])
    assert np.array_equal(np.array(K.pad_segment(rows[:1500], last=True)), g['c%d_pad_last' % ci])
    assert np.array_equal(K.pack_rows(np.array(rows, dtype=np.int64)), g['c%d_split' % ci])


def test_bar_limit_windows_and_task_outputs():
    """convert.py:420-508 (F's segmentation; not executable without its file I/O, so checked against its stated rules)"""
    g = load_golden('codec')
    rows = [tuple(r) for r in g['c2_enc_pretrain'].tolist()]          # 300 bars: crosses the 255-bar limit once
    parts = K.split_by_bar_limit(rows)
    assert len(parts) == 2 and all(p[-1] == K.EOS_ROW for p in parts)
    assert sum(len(p) - 1 for p in parts) == len(rows)
    assert max(r[0] for r in parts[0][:-1]) <= 255 and min(r[0] for r in parts[1][:-1]) == 0
    first_late = next(r for r in rows if r[0] > 255)
    assert parts[1][0] == (first_late[0] - 256,) + first_late[1:]
    segs = K.segments_for_task(rows, 'pretrain')
    assert all(len(s) == 1024 for s in segs)
    pairs = K.segments_for_task(rows, 'generate')
    for prompt, cont in pairs:
        assert len(prompt) == 1024 and len(cont) == 1024
        assert sum(1 for r in prompt if r[0] == 259) == 1
        body = [r for r in prompt if r[0] < 256]
        first_cont = cont[0]
        assert first_cont[0] >= max(r[0] for r in body)               # cut at a bar boundary: the continuation starts a new bar
    vel = K.segments_for_task([tuple(r) for r in g['c0_enc_velocity'].tolist()], 'velocity')
    seg, labels = vel[0]
    assert len(seg) == len(labels) and all(len(r) == 8 for r in seg) and labels[-1] == K.VELOCITY_MAP['OTHER']
    blocks = K.pack_rows(np.array(labels), K.VELOCITY_MAP['OTHER'], 1)
    assert blocks.shape == (1, 1024, 1) and blocks[0, len(labels):, 0].tolist() == [6] * (1024 - len(labels))


def test_midi_file_round_trip(tmp_path):
    g = load_golden('codec')
    for ci in (0, 1):
        sc = _score(g, ci)
        want = K.score_to_octuple(sc)
        decoded = K.octuple_to_score([r for r in want if r[2] <= 128])       # (the reference's decoder drops program 129)
        K.write_midi(decoded, str(tmp_path / 'a.mid'))
        again = K.read_midi(str(tmp_path / 'a.mid'))
        assert again.ticks_per_beat == 480
        # a decoded score is already on the codec's grid: onsets, programs, pitches, velocities and time signatures survive the
        # file exactly.  (Durations do not always: when two notes of one pitch overlap on a channel a MIDI file cannot say
        # which note-off ends which note - the reader closes every open note of the pitch at the first note-off, like miditoolkit.)
        e1, e2 = K.score_to_octuple(decoded), K.score_to_octuple(again)
        assert [r[:4] + r[5:7] for r in e1] == [r[:4] + r[5:7] for r in e2]
        assert max(abs(a[7] - b[7]) for a, b in zip(e1, e2)) <= 1              # tempo passes through microseconds per beat
    # without same-pitch overlaps the round trip is exact, durations included
    sc = K.Score(480, [K.Instrument(5, False, 'PIANO', [K.Note(30 * i, 30 * i + 30 * (1 + i % 7), 40 + i % 40, 4 * (i % 30) + 2)
                                                        for i in range(0, 400, 3)]),
                       K.Instrument(0, True, 'drums', [K.Note(120 * i, 120 * i + 60, 36 + i % 3, 102) for i in range(60)])],
                  [K.TimeSignature(3, 4, 0), K.TimeSignature(4, 4, 3 * 480 * 4)], [K.TempoChange(code_t, t) for code_t, t in
                                                                                 ((K.code_to_tempo(30), 0), (K.code_to_tempo(41), 960))])
    K.write_midi(sc, str(tmp_path / 'b.mid'))
    again = K.read_midi(str(tmp_path / 'b.mid'))
    assert K.score_to_octuple(again) == K.score_to_octuple(sc)
    assert [i.is_drum for i in again.instruments] == [False, True] and again.instruments[0].program == 5


def _write_synthetic_midis(root, n, seed):
    rs = np.random.RandomState(seed)
    paths = []
    for k in range(n):
        notes = [K.Note(int(t), int(t) + int(rs.choice([60, 120, 240, 480])), int(rs.randint(40, 90)), int(rs.randint(20, 120)))
                 for t in np.sort(rs.randint(0, 480 * 4 * (8 + 3 * k), size=120 + 40 * k)) // 30 * 30]
        sc = K.Score(480, [K.Instrument(int(rs.randint(0, 100)), False, 'PIANO', notes)], [K.TimeSignature(4, 4, 0)],
                     [K.TempoChange(float(rs.choice([90.0, 120.0, 150.0])), 0)])
        p = root / ('Q%d_%02d.mid' % (1 + k % 4, k))
        K.write_midi(sc, str(p))
        paths.append(p)
    return paths


@pytest.mark.parametrize('task', ['pretrain', 'generate', 'velocity', 'emotion'])
def test_convert_entry_point_builds_dataset_blocks(task, tmp_path):
    """main.convert() mirror of the reference dataset builder (convert.py:583-650): files -> `<dataset>_<split>.npy` blocks"""
    from pianobart_b200 import main as M
    src = tmp_path / 'midi'
    src.mkdir()
    _write_synthetic_midis(src, 10, 3)
    stats = M.convert(['--input', str(src), '--output', str(tmp_path / 'out'), '--dataset', 'syn', '--task', task, '--seed', '1'])
    assert sum(v[0] for v in stats.values()) == 10 and stats['train'][0] == 8
    x = np.load(tmp_path / 'out' / 'syn_train.npy')
    assert x.ndim == 3 and x.shape[1:] == (1024, 8)
    assert (x[..., 0] <= 259).all() and ((x[:, :, 0] == 256) | (x[:, :, 0] == 259) | (x[:, :, 0] < 256)).all()
    if task == 'pretrain':
        assert x.shape[0] == 8 and (x[:, -1] == np.array(K.PAD_ROW)).all()
        first = [tuple(r) for r in x[0].tolist() if r[0] < 256]
        assert first == sorted(first)                              # notes of a piece are sorted (bar, pos, program, pitch, ...)
    else:
        y = np.load(tmp_path / 'out' / 'syn_train_ans.npy')
        if task == 'generate':
            assert y.shape == x.shape
        elif task == 'velocity':
            assert y.shape == x.shape[:2] + (1,) and y.max() <= 6
        else:
            assert y.shape == (x.shape[0],) and set(y.tolist()) <= {0, 1, 2, 3}

def _smf(fmt, div, tracks):
    import struct
    return b'MThd' + struct.pack('>IHHH', 6, fmt, len(tracks), div) + b''.join(
        b'MTrk' + struct.pack('>I', len(t)) + t for t in tracks)


def test_midi_reader_on_hand_built_files(tmp_path):
    """Bytes written by hand (not by write_midi): running status, note-on velocity 0 as note-off, multi-byte delta times, sysex
    and unknown meta events, programs, channel 10, and the miditoolkit / pretty_midi pairing rule for same-pitch overlaps."""
    conductor = bytes([0x00, 0xff, 0x58, 0x04, 3, 2, 24, 8,                       # 3/4 at tick 0
                       0x00, 0xff, 0x51, 0x03, 0x07, 0xa1, 0x20,                  # 500000 us per beat = 120 bpm
                       0x00, 0xff, 0x7f, 0x02, 0x01, 0x02,                        # sequencer-specific meta: skipped
                       0x87, 0x40, 0xff, 0x51, 0x03, 0x0f, 0x42, 0x40,            # delta 960 (two-byte varlen): 60 bpm
                       0x00, 0xff, 0x2f, 0x00])
    piano = bytes([0x00, 0xff, 0x03, 0x05]) + b'Right' + bytes([
        0x00, 0xc0, 0x05,                                                          # program 5 on channel 0
        0x00, 0xf0, 0x03, 0x7e, 0x7f, 0xf7,                                        # sysex: skipped
        0x00, 0x90, 60, 100,                                                       # t=0    C4 on
        0x00, 64, 90,                                                              #        E4 on   (running status)
        0x78, 60, 0,                                                               # t=120  C4 off as note-on velocity 0 (running)
        0x00, 60, 80,                                                              # t=120  C4 on again on the same tick
        0x3c, 60, 70,                                                              # t=180  C4 on: overlaps the previous C4
        0x3c, 0x80, 60, 0,                                                         # t=240  ONE note-off closes BOTH open C4s
        0x00, 64, 0,                                                               # t=240  E4 off (running status 0x80)
        0x00, 0xc0, 0x07,                                                          # program change to 7 ...
        0x00, 0x90, 67, 50,                                                        # t=240  G4 on
        0x00, 0xc0, 0x09,                                                          # ... and to 9 before its note-off
        0x1e, 0x80, 67, 0,                                                         # t=270  G4 off: the note belongs to program 9
        0x00, 0x80, 72, 0,                                                         # spurious note-off: ignored
        0x00, 0xff, 0x2f, 0x00])
    drums = bytes([0x00, 0x99, 36, 110, 0x1e, 0x89, 36, 0, 0x00, 0xff, 0x2f, 0x00])   # channel 10
    path = str(tmp_path / 'hand.mid')
    open(path, 'wb').write(_smf(1, 480, [conductor, piano, drums]))
    sc = K.read_midi(path)
    assert sc.ticks_per_beat == 480
    assert [(c.numerator, c.denominator, c.time) for c in sc.time_signature_changes] == [(3, 4, 0)]
    assert [(round(c.tempo, 6), c.time) for c in sc.tempo_changes] == [(120.0, 0), (60.0, 960)]
    assert [(i.program, i.is_drum, i.name) for i in sc.instruments] == [(5, False, 'Right'), (9, False, 'Right'), (0, True, '')]
    notes = lambda inst: [(n.start, n.end, n.pitch, n.velocity) for n in inst.notes]
    assert notes(sc.instruments[0]) == [(0, 120, 60, 100), (120, 240, 60, 80), (180, 240, 60, 70), (0, 240, 64, 90)]
    assert notes(sc.instruments[1]) == [(240, 270, 67, 50)]
    assert notes(sc.instruments[2]) == [(0, 30, 36, 110)]
    # format 0: everything in one track
    open(path, 'wb').write(_smf(0, 96, [conductor[:-4] + piano[9:]]))
    sc0 = K.read_midi(path)
    assert sc0.ticks_per_beat == 96 and len(sc0.tempo_changes) == 2 and sum(len(i.notes) for i in sc0.instruments) == 5
    # the reader feeds the encoder: one Octuple row per note, programs kept (drums: max_inst + 1 = 129, convert.py:214)
    rows = K.score_to_octuple(sc)
    assert len(rows) == 6 and sorted({r[2] for r in rows}) == [5, 9, 129]
    with pytest.raises(ValueError):
        open(path, 'wb').write(b'RIFF' + bytes(20))
        K.read_midi(path)
