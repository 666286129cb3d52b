"""Hand-off from Extract to Decode without the JSON round trip.

In the reference, stage 1 writes ``extract.json`` (extractor.py:432-446) and stage 3 reads it back to build the condition
events of the decoder: ``TinyREMITokenizer(tempo_path).encode(extract.json)`` (infer.py:180-181, tokenizer.py:231-297).
``condition_events`` produces the same event sequence directly from the note records the device note stage returns
(the structured array of ``extract_many(..., as_dicts=False)`` or the list of dicts of ``_mpe2note``): the
``min_duration`` filter of ``_note2json``, the measure grid of ``_create_measures`` (tokenizer.py:174-229), the
position / duration quantisation of ``_assign_notes`` (231-255; ``_compute_rel_pos`` 143-160 with
``allow_triplet=False``, ``_map_duration_to_token_value`` 125-140) and the per-position pitch ordering / de-duplication
of ``encode`` (265-297).  Events are ``(type, value)`` tuples -- ``str(Event)`` in the reference is ``f"{type}_{value}"``
(vocab.py:21-36), see ``event_tokens``.

Host-side integer / float64 bookkeeping on a few thousand notes per song: there is nothing here for the GPU to do.
Arithmetic follows the reference expression by expression (Python floats), so the sequence is identical, not just close.
"""
import bisect
import json

_ALLOWED_DURATIONS_IN_16THS = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32)          # tokenizer.py:19
_REL_POS = ((0, 0), (1 / 4, 2), (1 / 2, 4), (3 / 4, 6), (1, 8))           # rel_pos_2_idx without triplets (tokenizer.py:144)


def load_tempo(tempo):
    """``tempo``: path of a tempo.json (beat_analyzer output) or the already parsed list of regions."""
    if isinstance(tempo, (list, tuple)):
        return list(tempo)
    with open(tempo, "r") as f:
        return json.load(f)


def measures_of(tempo_data):
    """tokenizer.py:174-229: one measure per downbeat (ending at the next downbeat / the next region's start / one bar
    later for the very last one), plus one synthetic bar before the first downbeat and one after the last measure.
    Returns a list of dicts {bpm, start, end, time_sig}."""
    out = []
    n_regions = len(tempo_data)
    for ri, region in enumerate(tempo_data):
        downbeats = region.get("downbeats", [])
        if not downbeats:
            continue
        bpm, ts = region["bpm"], region["time_sig"]
        bar = ts * (60 / bpm)
        nxt = tempo_data[ri + 1]["start"] if ri < n_regions - 1 else None
        for i, start in enumerate(downbeats):
            if i < len(downbeats) - 1:
                end = downbeats[i + 1]
            elif nxt is not None:
                end = nxt
            else:
                end = start + bar
            out.append({"bpm": bpm, "start": start, "end": end, "time_sig": ts})
    first, last = tempo_data[0], tempo_data[-1]
    fd = first["downbeats"][0]
    fbar = (60 / first["bpm"]) * first["time_sig"]
    out.insert(0, {"bpm": first["bpm"], "start": fd - fbar, "end": fd, "time_sig": first["time_sig"]})
    ld = last["downbeats"][-1]
    lbar = (60 / last["bpm"]) * last["time_sig"]
    out.append({"bpm": last["bpm"], "start": ld + lbar, "end": ld + 2 * lbar, "time_sig": last["time_sig"]})
    return out


def _rel_pos(onset, start, end, ts):
    """tokenizer.py:143-160 with allow_triplet=False."""
    m_rel = max(0.0, min(1.0, (onset - start) / (end - start)))
    b_idx = int(m_rel / (1 / ts))
    b_rel = (m_rel % (1 / ts)) / (1 / ts)
    best = min(_REL_POS, key=lambda kv: abs(kv[0] - b_rel))[1]            # first minimum wins, like min() over the dict keys
    pos = b_idx * 8 + best
    return pos, pos >= 8 * ts


def _duration_token(duration_sec, bpm):
    """tokenizer.py:125-140."""
    if duration_sec <= 0 or bpm <= 0:
        return _ALLOWED_DURATIONS_IN_16THS[0]
    per16 = (60.0 / bpm) / 4.0
    d = duration_sec / per16
    return min(_ALLOWED_DURATIONS_IN_16THS, key=lambda x: abs(x - d))


def _as_columns(notes):
    """(onset, offset, pitch) python lists from a structured array or a list of dicts, in the given order."""
    if hasattr(notes, "dtype") and notes.dtype.names:
        return notes["onset"].tolist(), notes["offset"].tolist(), [int(p) for p in notes["pitch"].tolist()]
    return [float(n["onset"]) for n in notes], [float(n["offset"]) for n in notes], [int(n["pitch"]) for n in notes]


def condition_events(notes, tempo, min_duration=0.0):
    """Note records (in ``_mpe2note`` order) -> the condition event sequence of the decoder, as ``(type, value)`` tuples.

    ``min_duration`` is ``config.infer.min_duration`` (0.08): ``_note2json`` drops shorter notes before the tokenizer ever
    sees them (extractor.py:438-440).  ``tempo``: path or parsed content of tempo.json."""
    measures = measures_of(load_tempo(tempo))
    onset, offset, pitch = _as_columns(notes)
    starts = [m["start"] for m in measures]
    ordered = all(measures[i]["end"] <= measures[i + 1]["start"] for i in range(len(measures) - 1)) and starts == sorted(starts)
    chords = [dict() for _ in measures]          # measure -> {pos_idx: [(pitch, duration)] in arrival order}

    def find(t):                                 # first measure in list order with start <= t < end (tokenizer.py:234-235)
        if ordered:
            k = bisect.bisect_right(starts, t) - 1
            return k if k >= 0 and t < measures[k]["end"] else -1
        for k, m in enumerate(measures):
            if m["start"] <= t < m["end"]:
                return k
        return -1

    for on, off, p in zip(onset, offset, pitch):
        if off - on < min_duration:
            continue
        k = find(on)
        if k < 0:
            continue
        m = measures[k]
        pos, is_last = _rel_pos(on, m["start"], m["end"], m["time_sig"])
        dur = _duration_token(off - on, m["bpm"])
        if is_last and k + 1 < len(measures):
            chords[k + 1].setdefault(0, []).append((p, dur))
        elif not is_last:
            chords[k].setdefault(pos, []).append((p, dur))
    events = []
    for ch in chords:
        events.append(("Bar", "BOS"))
        for pos in sorted(ch):
            events.append(("Pos", pos))
            seen = set()
            for p, dur in sorted(ch[pos], key=lambda x: -x[0]):      # stable: the first note of a pitch wins (tokenizer.py:283-285)
                if p in seen:
                    continue
                seen.add(p)
                events.append(("Note", p))
                events.append(("Duration", dur))
        events.append(("Bar", "EOS"))
    return events


def event_tokens(events):
    """``(type, value)`` tuples -> the reference's token strings (``str(Event)``, vocab.py:33-35)."""
    return [f"{t}_{v}" for t, v in events]
