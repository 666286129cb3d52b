// Piano-rolls -> notes on the device: one thread per (song, pitch) walks the time axis once.
//
// Replaces AMTAPC_Extractor._mpe2note (reference etude/data/extractor.py:256-418), bit-exactly on identical rolls:
//  * a frame is an onset/offset peak iff value >= threshold and, skipping equal neighbours outward, the first
//    different value on each side is smaller (270-286, 299-316)  ->  evaluated per maximal run of equal values;
//  * peak times use the reference's NumPy-2 (NEP 50) arithmetic: interpolated times are float32
//    f32(i*hop) -/+ f32(f32(hop/2) * (a-b)) / (c-d), everything else is float64 k*hop.
//    __fmul_rn/__fdiv_rn/__fsub_rn/__fadd_rn keep nvcc from contracting or reassociating them;
//  * note end = f(next onset, first offset peak after the onset, first frame with mpe < thr) per 331-404;
//  * velocity 0 is dropped in 'ignore_zero' mode; an overlapping previous note of the pitch is clipped (411-414).
// Two passes (count, then fill at an exclusive-scan offset) give an exact-size, pitch-major note array per song;
// the final stable sort by onset (416) is done by the host wrapper.
#pragma once
#include "common.cuh"

namespace etude {

struct NoteRec {
    int32_t pitch;
    int32_t velocity;
    double onset;
    double offset;
};

struct NotesSong {
    int64_t row_off;  // first roll row of the song
    int64_t n_rows;   // T_pad
};

struct NotesParams {
    const float* onset;
    const float* offset;
    const float* mpe;
    const int8_t* velocity;
    const NotesSong* songs;
    int n_songs;
    int note_min;
    double hop_sec;
    float thr_onset, thr_offset, thr_mpe;
    int mode_velocity;  // 0 ignore_zero, 1 org
    int mode_offset;    // 0 shorter, 1 longer, 2 offset
    int64_t* counts;        // [n_songs * 88]
    const int64_t* starts;  // [n_songs * 88] exclusive scan of counts (fill pass)
    NoteRec* notes;         // fill pass output
};

struct PeakIter {
    const float* a;  // column base, stride kNotes
    int64_t T;
    float thr;
    int64_t pos;      // next frame to examine
    int64_t run_end;  // last frame of the current qualifying run
    bool in_run;
};

__device__ __forceinline__ float roll_at(const float* a, int64_t i) { return __ldg(a + i * kNotes); }

// Advances to the next peak; returns false when the series is exhausted.
__device__ bool next_peak(PeakIter& it, double hop_sec, int64_t& loc, double& time) {
    for (;;) {
        if (it.in_run) {
            if (it.pos <= it.run_end) {
                const int64_t i = it.pos++;
                loc = i;
                const double ti = (double)i * hop_sec;
                if (i == 0 || i == it.T - 1) {
                    time = ti;
                } else {
                    const float v = roll_at(it.a, i), p = roll_at(it.a, i - 1), q = roll_at(it.a, i + 1);
                    const float half_hop = (float)(hop_sec * 0.5);
                    if (p == q) {
                        time = ti;
                    } else if (p > q) {
                        const float frac = __fdiv_rn(__fmul_rn(half_hop, __fsub_rn(p, q)), __fsub_rn(v, q));
                        time = (double)__fsub_rn((float)ti, frac);
                    } else {
                        const float frac = __fdiv_rn(__fmul_rn(half_hop, __fsub_rn(q, p)), __fsub_rn(v, p));
                        time = (double)__fadd_rn((float)ti, frac);
                    }
                }
                return true;
            }
            it.in_run = false;
        }
        if (it.pos >= it.T) return false;
        const int64_t s = it.pos;
        const float v = roll_at(it.a, s);
        int64_t e = s;
        while (e + 1 < it.T && roll_at(it.a, e + 1) == v) ++e;
        it.pos = e + 1;
        if (v >= it.thr) {
            const bool left = (s == 0) || (v > roll_at(it.a, s - 1));
            const bool right = (e == it.T - 1) || (v > roll_at(it.a, e + 1));
            if (left && right) {
                it.in_run = true;
                it.pos = s;
                it.run_end = e;
            }
        }
    }
}

template <bool FILL>
__global__ void notes_kernel(const NotesParams p) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.n_songs * kNotes) return;
    const int song = gid / kNotes, j = gid % kNotes;
    const NotesSong sg = p.songs[song];
    const int64_t T = sg.n_rows;
    const float* on = p.onset + sg.row_off * kNotes + j;
    const float* off = p.offset + sg.row_off * kNotes + j;
    const float* mpe = p.mpe + sg.row_off * kNotes + j;
    const int8_t* vel = p.velocity + sg.row_off * kNotes + j;

    PeakIter it_on{on, T, p.thr_onset, 0, 0, false};
    PeakIter it_off{off, T, p.thr_offset, 0, 0, false};
    int64_t count = 0;
    NoteRec* dst = FILL ? p.notes + p.starts[gid] : nullptr;
    NoteRec prev{0, 0, 0.0, 0.0};  // last appended note of this pitch, not yet stored (its end may still be clipped)
    bool have_prev = false;

    int64_t loc_onset, loc_nextpk = 0, loc_offpk = -1;
    double time_onset, time_nextpk = 0.0, time_offpk = 0.0;
    bool have_cur = next_peak(it_on, p.hop_sec, loc_onset, time_onset);
    bool have_off = next_peak(it_off, p.hop_sec, loc_offpk, time_offpk);
    double time_offset = 0.0, time_mpe = 0.0;
    while (have_cur) {
        const bool have_next = next_peak(it_on, p.hop_sec, loc_nextpk, time_nextpk);
        int64_t loc_next;
        double time_next;
        if (have_next) {
            loc_next = loc_nextpk;
            time_next = time_nextpk;
        } else {
            loc_next = T;
            time_next = (double)(T - 1) * p.hop_sec;
        }
        // first offset peak strictly after the onset frame
        while (have_off && loc_offpk <= loc_onset) have_off = next_peak(it_off, p.hop_sec, loc_offpk, time_offpk);
        int64_t loc_offset = loc_onset + 1;
        bool flag_offset = false;
        if (have_off) {
            loc_offset = loc_offpk;
            time_offset = time_offpk;
            flag_offset = true;
        }
        if (loc_offset > loc_next) {
            loc_offset = loc_next;
            time_offset = time_next;
        }
        int64_t loc_mpe = loc_onset + 1;
        bool flag_mpe = false;
        for (int64_t ii = loc_onset + 1; ii < loc_next; ++ii) {
            if (__ldg(mpe + ii * kNotes) < p.thr_mpe) {
                loc_mpe = ii;
                flag_mpe = true;
                time_mpe = (double)ii * p.hop_sec;
                break;
            }
        }
        const int velocity_value = (int)__ldg(vel + loc_onset * kNotes);
        double offset_value;
        if (!flag_offset && !flag_mpe) offset_value = time_next;
        else if (flag_offset && !flag_mpe) offset_value = time_offset;
        else if (!flag_offset && flag_mpe) offset_value = time_mpe;
        else if (p.mode_offset == 2) offset_value = time_offset;
        else if (p.mode_offset == 1) offset_value = (loc_offset >= loc_mpe) ? time_offset : time_mpe;
        else offset_value = (loc_offset <= loc_mpe) ? time_offset : time_mpe;

        if (p.mode_velocity != 0 || velocity_value > 0) {
            if (have_prev) {
                if (time_onset < prev.offset) prev.offset = time_onset;  // extractor.py:411-414 (same pitch by construction)
                if (FILL) dst[count - 1] = prev;
            }
            prev.pitch = j + p.note_min;
            prev.velocity = velocity_value;
            prev.onset = time_onset;
            prev.offset = offset_value;
            have_prev = true;
            ++count;
        }
        have_cur = have_next;
        loc_onset = loc_nextpk;
        time_onset = time_nextpk;
    }
    if (FILL && have_prev) dst[count - 1] = prev;
    if (!FILL) p.counts[gid] = count;
}

}  // namespace etude
