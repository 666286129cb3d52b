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
// Two passes (count, then fill at an exclusive-scan offset) give an exact-size, pitch-major note array per song.
// The walk of one (song, pitch) is inherently serial and branchy, so it gets a warp to itself (lane 0 walks; 32
// different pitches in one warp would serialise on divergence).  The final ordering sorted(sorted(a, key=pitch),
// key=onset) (416) is also done here: every pitch's notes are already in onset order, so a note's final position is
// its rank, found with one binary search per other pitch (notes_rank_kernel), and the host receives sorted songs.
// The three fp32 rolls are first transposed to pitch-major [88][T] per song (notes_transpose_kernel) so that each
// thread's serial walk along time reads contiguous memory (L1 hits) instead of one 352-byte-strided load per frame.
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
    const float* onset;   // pitch-major copies: song s, pitch j, frame i at [row_off[s] * 88 + j * n_rows[s] + i]
    const float* offset;
    const float* mpe;
    const int8_t* velocity;  // frame-major [rows, 88] (read at onset frames only)
    const NotesSong* songs;
    int n_songs;
    int note_min;
    double hop_sec;
    float thr_onset, thr_offset, thr_mpe;
    int mode_velocity;  // 0 ignore_zero, 1 org
    int mode_offset;    // 0 shorter, 1 longer, 2 offset
    int64_t* counts;        // [n_songs * 88]
    const int64_t* starts;  // [n_songs * 88] exclusive scan of counts (fill pass)
    NoteRec* notes;         // fill pass output (pitch-major per song)
    double* onsets;         // fill pass output: onset of every note, same order (dense search keys of the rank pass)
};

struct PeakIter {
    const float* a;  // this pitch's series, contiguous in time
    int64_t T;
    float thr;
    int64_t pos;      // next frame to examine
    int64_t run_end;  // last frame of the current qualifying run
    bool in_run;
};

__device__ __forceinline__ float roll_at(const float* a, int64_t i) { return __ldg(a + i); }

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

// rolls [rows, 88] (frame-major) -> per song [88][n_rows] (pitch-major); grid (row tiles of 32, songs, 3 arrays)
__global__ void __launch_bounds__(256)
notes_transpose_kernel(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ a2,
                       const NotesSong* __restrict__ songs, float* __restrict__ t0, float* __restrict__ t1, float* __restrict__ t2) {
    __shared__ float tile[32][33];
    const NotesSong sg = songs[blockIdx.y];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    if (r0 >= sg.n_rows) return;
    const float* src = (blockIdx.z == 0 ? a0 : (blockIdx.z == 1 ? a1 : a2)) + sg.row_off * kNotes;
    float* dst = (blockIdx.z == 0 ? t0 : (blockIdx.z == 1 ? t1 : t2)) + sg.row_off * kNotes;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int c0 = 0; c0 < kNotes; c0 += 32) {
        for (int rr = ty; rr < 32; rr += 8) {
            const int64_t r = r0 + rr;
            const int c = c0 + tx;
            tile[rr][tx] = (r < sg.n_rows && c < kNotes) ? __ldg(src + r * kNotes + c) : 0.f;
        }
        __syncthreads();
        for (int cc = ty; cc < 32; cc += 8) {
            const int c = c0 + cc;
            const int64_t r = r0 + tx;
            if (c < kNotes && r < sg.n_rows) dst[(int64_t)c * sg.n_rows + r] = tile[tx][cc];
        }
        __syncthreads();
    }
}

// Appends the finished note `rec` as element `idx` of its pitch's list, keeping the list ordered by onset (stable).
// Peak times are monotone in the frame index except for plateau neighbours, whose interpolated times
// i*hop + hop/2 and (i+1)*hop - hop/2 can tie or invert by an ulp; the rank pass needs every pitch list sorted, and
// sorting (onset, emission order) inside a pitch first does not change the result of the reference's stable sort.
__device__ __forceinline__ void store_sorted(NoteRec* dst, double* dst_on, int64_t idx, const NoteRec& rec) {
    int64_t k = idx;
    while (k > 0 && dst_on[k - 1] > rec.onset) {
        dst[k] = dst[k - 1];
        dst_on[k] = dst_on[k - 1];
        --k;
    }
    dst[k] = rec;
    dst_on[k] = rec.onset;
}

template <bool FILL>
__device__ void notes_walk(const NotesParams& p, const int gid) {
    const int song = gid / kNotes, j = gid % kNotes;
    const NotesSong sg = p.songs[song];
    const int64_t T = sg.n_rows;
    const float* on = p.onset + sg.row_off * kNotes + (int64_t)j * T;
    const float* off = p.offset + sg.row_off * kNotes + (int64_t)j * T;
    const float* mpe = p.mpe + sg.row_off * kNotes + (int64_t)j * T;
    const int8_t* vel = p.velocity + sg.row_off * kNotes + j;

    PeakIter it_on{on, T, p.thr_onset, 0, 0, false};
    PeakIter it_off{off, T, p.thr_offset, 0, 0, false};
    int64_t count = 0;
    NoteRec* dst = FILL ? p.notes + p.starts[gid] : nullptr;
    double* dst_on = FILL ? p.onsets + p.starts[gid] : nullptr;
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
            if (__ldg(mpe + ii) < p.thr_mpe) {
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
                if (FILL) store_sorted(dst, dst_on, count - 1, prev);
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
    if (FILL && have_prev) store_sorted(dst, dst_on, count - 1, prev);
    p.counts[gid] = count;
}

// One-warp blocks, at most one per SM (grid <= number of SMs, every block walks (song, pitch) items grid-stride; lane 0
// walks, 32 different pitches in one warp would serialise on divergence).  In extract_many these kernels run on the notes
// stream while the persistent model kernels of the next song group run on the launching stream: those leave ~4 K
// registers and ~2 KB of shared memory per SM, enough for ONE such block.  More blocks per SM (which the scheduler
// happily places while the SMs are empty between two model kernels) keep the next model kernel's CTAs off their SMs until
// the walks finish -- milliseconds on a kernel that is statically partitioned over all SMs.
template <bool FILL>
__global__ void __launch_bounds__(32) notes_kernel(const NotesParams p) {
    if (threadIdx.x != 0) return;
    for (int gid = blockIdx.x; gid < p.n_songs * kNotes; gid += gridDim.x) notes_walk<FILL>(p, gid);
}

// Final order of a song's notes: stable sort by onset of the pitch-major array (extractor.py:416).  Note k of pitch j
// lands at  k + sum over pitches j' != j of #{notes of j' with onset < t, or onset == t and j' < j}.
// The walk pass leaves every (song, pitch) run in its own slab (slab_base[song * 88 + j], counts[...] entries, capacity
// = the song's frame count: at most one note per frame), so no count pass and no host round trip precede the fill.
// grid (ceil(max notes per song / blockDim), n_songs); out_base[song] = first record of the song in `sorted`.
__global__ void __launch_bounds__(256)
notes_rank_kernel(const NoteRec* __restrict__ notes, const double* __restrict__ onsets, const int64_t* __restrict__ slab_base,
                  const int64_t* __restrict__ counts, const int64_t* __restrict__ out_base, NoteRec* __restrict__ sorted) {
    __shared__ int64_t s_base[kNotes];
    __shared__ int64_t s_pref[kNotes + 1];   // dense prefix of the song's per-pitch counts
    const int song = blockIdx.y;
    for (int j = threadIdx.x; j < kNotes; j += blockDim.x) s_base[j] = slab_base[song * kNotes + j];
    if (threadIdx.x == 0) {
        int64_t acc = 0;
        for (int j = 0; j < kNotes; ++j) { s_pref[j] = acc; acc += counts[song * kNotes + j]; }
        s_pref[kNotes] = acc;
    }
    __syncthreads();
    const int64_t n = s_pref[kNotes];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = 0;  // pitch run containing dense index i: last j with s_pref[j] <= i
    {
        int lo = 0, hi = kNotes - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_pref[mid] <= i) lo = mid; else hi = mid - 1;
        }
        j = lo;
    }
    const int64_t k = i - s_pref[j];
    const int64_t g = s_base[j] + k;
    const double t = onsets[g];
    int64_t rank = k;
    for (int jj = 0; jj < kNotes; ++jj) {
        if (jj == j) continue;
        const int64_t first = s_base[jj];
        int64_t lo = first, hi = first + (s_pref[jj + 1] - s_pref[jj]);
        if (jj < j) {  // upper bound: first onset > t
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(onsets + mid) <= t) lo = mid + 1; else hi = mid;
            }
        } else {       // lower bound: first onset >= t
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(onsets + mid) < t) lo = mid + 1; else hi = mid;
            }
        }
        rank += lo - first;
    }
    sorted[out_base[song] + rank] = notes[g];
}

}  // namespace etude
