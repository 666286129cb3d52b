// Piano-rolls -> notes on the device, parallel over (song, pitch, 256-frame chunk).
//
// Replaces AMTAPC_Extractor._mpe2note (reference etude/data/extractor.py:256-418), bit-exactly on identical rolls:
//  * a frame is an onset/offset peak iff value >= threshold and, skipping equal neighbours outward, the first
//    different value on each side is smaller (270-286, 299-316)  ->  evaluated per maximal run of equal values;
//  * peak times use the reference's NumPy-2 (NEP 50) arithmetic: interpolated times are float32
//    f32(i*hop) -/+ f32(f32(hop/2) * (a-b)) / (c-d), everything else is float64 k*hop.
//    __fmul_rn/__fdiv_rn/__fsub_rn/__fadd_rn keep nvcc from contracting or reassociating them;
//  * note end = f(next onset, first offset peak after the onset, first frame with mpe < thr) per 331-404;
//  * velocity 0 is dropped in 'ignore_zero' mode; an overlapping previous note of the pitch is clipped (411-414).
//
// The reference walks every pitch serially along time; one thread per (song, pitch) did the same here in round 1 and cost
// ~8 ms per call whatever the number of songs.  Now the time axis is cut into chunks of 256 frames and one thread owns the
// notes whose ONSET lies in its chunk.  Everything a note needs beyond its chunk is a "first ... after" query:
//   the next onset peak (its frame bounds the note, its time may end it), the first offset peak after the onset,
//   the next KEPT note of the pitch (velocity > 0 in 'ignore_zero' mode: its onset clips this note's end, 411-414);
// a pre-pass (notes_scan_kernel) records, per chunk, the first onset peak, the first kept onset peak and the first offset
// peak, so those queries cost one look at a handful of chunk entries instead of a walk to the end of the song (offset peaks
// are rare with offset_threshold = 1.0: the serial walk for "first offset after" is what a naive chunking would repeat per
// chunk).  Only the mpe scan (first frame below threshold between this onset and the next) stays a frame walk, bounded by the
// next onset exactly as in the reference.
// Passes: transpose rolls to pitch-major (contiguous walks) -> scan -> walk (notes into per-chunk slabs: at most one note
// per frame) -> compact per (song, pitch) in place + counts -> per-song totals / bases -> rank (the final ordering
// sorted(sorted(a, key=pitch), key=onset) (416): a note's position is its rank, one binary search per other pitch).
// No pass needs a host round trip; the host reads the per-song counts once, after the last kernel.
#pragma once
#include "common.cuh"

namespace etude {

constexpr int kNoteChunk = 256;   // frames per (song, pitch) work item

struct NoteRec {
    int32_t pitch;
    int32_t velocity;
    double onset;
    double offset;
};

struct NotesSong {
    int64_t row_off;    // first roll row of the song (frame-major rolls)
    int64_t n_rows;     // T_pad
    int64_t t_off;      // first row of the song in the pitch-major scratch / first slab record = t_off * 88
    int32_t chunk_off;  // first entry of the song in the per-chunk arrays (88 * n_chunks entries per song, pitch-major)
    int32_t n_chunks;   // ceil(n_rows / kNoteChunk)
};

struct NotesParams {
    const float* onset;   // pitch-major copies: song s, pitch j, frame i at [t_off[s] * 88 + j * n_rows[s] + i]
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
    int32_t* first_on;       // per chunk: frame of the first onset peak in the chunk, or -1
    int32_t* first_kept;     // per chunk: frame of the first onset peak that yields a note (velocity rule), or -1
    int32_t* first_off;      // per chunk: frame of the first offset peak in the chunk, or -1
    int32_t* chunk_count;    // per chunk: notes emitted by the walk
    int64_t* counts;         // [n_songs * 88] notes per (song, pitch) after compaction
    NoteRec* notes;          // slabs: (song, pitch) at t_off * 88 + j * n_rows, chunk c at + c * kNoteChunk; dense after compaction
    double* onsets;          // same layout: onset of every note (dense search keys of the rank pass)
};

// The per-item logic below is __host__ __device__: tests/notes_host.cu compiles it for the CPU and the CPU test suite
// checks the chunked algorithm against the reference's own results without a GPU (tests/test_notes_host.py).
#ifdef __CUDA_ARCH__
#define NOTE_LD(ptr) __ldg(ptr)
#define NOTE_FMUL(a, b) __fmul_rn(a, b)
#define NOTE_FDIV(a, b) __fdiv_rn(a, b)
#define NOTE_FSUB(a, b) __fsub_rn(a, b)
#define NOTE_FADD(a, b) __fadd_rn(a, b)
#else   // host build: compiled with -ffp-contract=off, so plain float arithmetic is the same IEEE sequence
#define NOTE_LD(ptr) (*(ptr))
#define NOTE_FMUL(a, b) ((a) * (b))
#define NOTE_FDIV(a, b) ((a) / (b))
#define NOTE_FSUB(a, b) ((a) - (b))
#define NOTE_FADD(a, b) ((a) + (b))
#endif
#define NOTE_HD __host__ __device__ __forceinline__

NOTE_HD float roll_at(const float* a, int64_t i) { return NOTE_LD(a + i); }

// Time of the peak at frame i (extractor.py:287-295 / 318-326).
NOTE_HD double peak_time(const float* a, int64_t i, int64_t T, double hop_sec) {
    const double ti = (double)i * hop_sec;
    if (i == 0 || i == T - 1) return ti;
    const float v = roll_at(a, i), p = roll_at(a, i - 1), q = roll_at(a, i + 1);
    const float half_hop = (float)(hop_sec * 0.5);
    if (p == q) return ti;
    if (p > q) {
        const float frac = NOTE_FDIV(NOTE_FMUL(half_hop, NOTE_FSUB(p, q)), NOTE_FSUB(v, q));
        return (double)NOTE_FSUB((float)ti, frac);
    }
    const float frac = NOTE_FDIV(NOTE_FMUL(half_hop, NOTE_FSUB(q, p)), NOTE_FSUB(v, p));
    return (double)NOTE_FADD((float)ti, frac);
}

// Iterates the peak FRAMES of one series in order, run by run (every frame of a qualifying plateau is a peak).
struct PeakIter {
    const float* a;  // this pitch's series, contiguous in time
    int64_t T;
    float thr;
    int64_t pos;      // next frame to examine
    int64_t run_end;  // last frame of the current qualifying run
    bool in_run;
};

// Starts at the run that contains frame `from` (its earlier frames are skipped by the caller).
NOTE_HD PeakIter peak_iter_at(const float* a, int64_t T, float thr, int64_t from) {
    int64_t s = from;
    const float v = roll_at(a, from);
    while (s > 0 && roll_at(a, s - 1) == v) --s;
    return PeakIter{a, T, thr, s, 0, false};
}

// Advances to the next peak frame < limit; returns false when there is none (the iterator may then be discarded).
__host__ __device__ inline bool next_peak(PeakIter& it, int64_t limit, int64_t& loc) {
    for (;;) {
        if (it.in_run) {
            if (it.pos <= it.run_end) {
                if (it.pos >= limit) return false;
                loc = it.pos++;
                return true;
            }
            it.in_run = false;
        }
        if (it.pos >= it.T || it.pos >= limit) return false;
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
    float* dst = (blockIdx.z == 0 ? t0 : (blockIdx.z == 1 ? t1 : t2)) + sg.t_off * kNotes;
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

// (song, pitch, chunk) of a thread: grid.y = song, threads along x cover pitch * n_chunks + chunk (consecutive lanes =
// consecutive chunks of one pitch)
__device__ __forceinline__ bool note_item(const NotesParams& p, NotesSong& sg, int& j, int& c) {
    sg = p.songs[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kNotes * sg.n_chunks) return false;
    j = idx / sg.n_chunks;
    c = idx - j * sg.n_chunks;
    return true;
}

// Pre-pass: first onset peak / first kept onset peak / first offset peak of every chunk.
__host__ __device__ inline void notes_scan_item(const NotesParams& p, const NotesSong& sg, int j, int c) {
    const int64_t T = sg.n_rows, c0 = (int64_t)c * kNoteChunk, c1 = (c0 + kNoteChunk < T) ? c0 + kNoteChunk : T;
    const float* on = p.onset + sg.t_off * kNotes + (int64_t)j * T;
    const float* off = p.offset + sg.t_off * kNotes + (int64_t)j * T;
    const int8_t* vel = p.velocity + sg.row_off * kNotes + j;
    int32_t f_on = -1, f_kept = -1, f_off = -1;
    int64_t loc;
    PeakIter it = peak_iter_at(on, T, p.thr_onset, c0);
    while (next_peak(it, c1, loc)) {
        if (loc < c0) continue;
        if (f_on < 0) f_on = (int32_t)loc;
        if (p.mode_velocity != 0 || NOTE_LD(vel + loc * kNotes) > 0) { f_kept = (int32_t)loc; break; }
    }
    it = peak_iter_at(off, T, p.thr_offset, c0);
    while (next_peak(it, c1, loc)) {
        if (loc < c0) continue;
        f_off = (int32_t)loc;
        break;
    }
    const int e = sg.chunk_off + j * sg.n_chunks + c;
    p.first_on[e] = f_on;
    p.first_kept[e] = f_kept;
    p.first_off[e] = f_off;
}
__global__ void __launch_bounds__(64) notes_scan_kernel(const NotesParams p) {
    NotesSong sg;
    int j, c;
    if (note_item(p, sg, j, c)) notes_scan_item(p, sg, j, c);
}

// First non-negative entry of `first` in chunks (c, n_chunks) of this pitch, or -1.
NOTE_HD int64_t first_after_chunk(const int32_t* first, int c, int n_chunks) {
    for (int k = c + 1; k < n_chunks; ++k) {
        const int32_t v = NOTE_LD(first + k);
        if (v >= 0) return v;
    }
    return -1;
}

// Appends the finished note `rec` as element `idx` of its chunk's list, keeping the list ordered by onset (stable).
// Peak times are monotone in the frame index except for plateau neighbours, whose interpolated times
// i*hop + hop/2 and (i+1)*hop - hop/2 can tie or invert by an ulp; the rank pass needs every pitch list sorted, and
// sorting (onset, emission order) inside a pitch first does not change the result of the reference's stable sort.
NOTE_HD void store_sorted(NoteRec* dst, double* dst_on, int64_t idx, const NoteRec& rec) {
    int64_t k = idx;
    while (k > 0 && dst_on[k - 1] > rec.onset) {
        dst[k] = dst[k - 1];
        dst_on[k] = dst_on[k - 1];
        --k;
    }
    dst[k] = rec;
    dst_on[k] = rec.onset;
}

// The walk: one thread emits the notes whose onset peak lies in its chunk (extractor.py:331-414).
__host__ __device__ inline void notes_walk_item(const NotesParams& p, const NotesSong& sg, int j, int c) {
    const int64_t T = sg.n_rows, c0 = (int64_t)c * kNoteChunk, c1 = (c0 + kNoteChunk < T) ? c0 + kNoteChunk : T;
    const int ce = sg.chunk_off + j * sg.n_chunks;          // chunk entries of this pitch
    if (NOTE_LD(p.first_on + ce + c) < 0) {                 // no onset peak in this chunk
        p.chunk_count[ce + c] = 0;
        return;
    }
    const float* on = p.onset + sg.t_off * kNotes + (int64_t)j * T;
    const float* off = p.offset + sg.t_off * kNotes + (int64_t)j * T;
    const float* mpe = p.mpe + sg.t_off * kNotes + (int64_t)j * T;
    const int8_t* vel = p.velocity + sg.row_off * kNotes + j;
    NoteRec* dst = p.notes + sg.t_off * kNotes + (int64_t)j * T + c0;
    double* dst_on = p.onsets + sg.t_off * kNotes + (int64_t)j * T + c0;

    // what lies beyond the chunk (one look at the chunk tables each)
    const int64_t far_on = first_after_chunk(p.first_on + ce, c, sg.n_chunks);
    const int64_t far_kept = first_after_chunk(p.first_kept + ce, c, sg.n_chunks);
    const int64_t far_off = first_after_chunk(p.first_off + ce, c, sg.n_chunks);
    const double far_on_time = far_on >= 0 ? peak_time(on, far_on, T, p.hop_sec) : 0.0;
    const double far_off_time = far_off >= 0 ? peak_time(off, far_off, T, p.hop_sec) : 0.0;

    PeakIter it_on = peak_iter_at(on, T, p.thr_onset, c0);
    PeakIter it_off = peak_iter_at(off, T, p.thr_offset, c0);
    int64_t loc_onset = -1, loc_nextpk = -1, loc_offpk = -1;
    bool have_cur = false, have_off = false;
    while ((have_cur = next_peak(it_on, c1, loc_onset)) && loc_onset < c0) {}
    while ((have_off = next_peak(it_off, c1, loc_offpk)) && loc_offpk < c0) {}
    double time_onset = have_cur ? peak_time(on, loc_onset, T, p.hop_sec) : 0.0;

    int64_t count = 0;
    NoteRec prev{0, 0, 0.0, 0.0};  // last appended note of this chunk, not yet stored (its end may still be clipped)
    bool have_prev = false;
    double time_offset = 0.0, time_mpe = 0.0;
    while (have_cur) {
        // the next onset peak: in this chunk, else the first one of a later chunk, else none
        const bool next_in = next_peak(it_on, c1, loc_nextpk);
        int64_t loc_next;
        double time_next, time_nextpk = 0.0;
        if (next_in) {
            loc_next = loc_nextpk;
            time_nextpk = peak_time(on, loc_nextpk, T, p.hop_sec);
            time_next = time_nextpk;
        } else if (far_on >= 0) {
            loc_next = far_on;
            time_next = far_on_time;
        } else {
            loc_next = T;
            time_next = (double)(T - 1) * p.hop_sec;
        }
        // first offset peak strictly after the onset frame: in this chunk, else the first one of a later chunk
        while (have_off && loc_offpk <= loc_onset) have_off = next_peak(it_off, c1, loc_offpk);
        int64_t loc_offset = loc_onset + 1;
        bool flag_offset = false;
        if (have_off) {
            loc_offset = loc_offpk;
            time_offset = peak_time(off, loc_offpk, T, p.hop_sec);
            flag_offset = true;
        } else if (far_off >= 0) {
            loc_offset = far_off;
            time_offset = far_off_time;
            flag_offset = true;
        }
        if (loc_offset > loc_next) {
            loc_offset = loc_next;
            time_offset = time_next;
        }
        int64_t loc_mpe = loc_onset + 1;
        bool flag_mpe = false;
        for (int64_t ii = loc_onset + 1; ii < loc_next; ++ii) {
            if (NOTE_LD(mpe + ii) < p.thr_mpe) {
                loc_mpe = ii;
                flag_mpe = true;
                time_mpe = (double)ii * p.hop_sec;
                break;
            }
        }
        const int velocity_value = (int)NOTE_LD(vel + loc_onset * kNotes);
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
                store_sorted(dst, dst_on, count - 1, prev);
            }
            prev.pitch = j + p.note_min;
            prev.velocity = velocity_value;
            prev.onset = time_onset;
            prev.offset = offset_value;
            have_prev = true;
            ++count;
        }
        have_cur = next_in;
        loc_onset = loc_nextpk;
        time_onset = time_nextpk;
    }
    if (have_prev) {
        if (far_kept >= 0) {   // the next note of this pitch starts in a later chunk: its onset may clip this one
            const double t = peak_time(on, far_kept, T, p.hop_sec);
            if (t < prev.offset) prev.offset = t;
        }
        store_sorted(dst, dst_on, count - 1, prev);
    }
    p.chunk_count[ce + c] = (int32_t)count;
}
__global__ void __launch_bounds__(64) notes_walk_kernel(const NotesParams p) {
    NotesSong sg;
    int j, c;
    if (note_item(p, sg, j, c)) notes_walk_item(p, sg, j, c);
}

// One warp per (song, pitch): pushes the chunk lists together, in place and in order (a chunk's dense position never
// exceeds its slab position, so a chunk read completely into registers before it is written cannot clobber unread data),
// repairs the order across chunk borders (the plateau-neighbour ulp inversions of store_sorted can straddle one) and writes
// the pitch's note count.
__global__ void __launch_bounds__(128) notes_compact_kernel(const NotesParams p) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= p.n_songs * kNotes) return;
    const int song = gw / kNotes, j = gw - song * kNotes;
    const NotesSong sg = p.songs[song];
    const int ce = sg.chunk_off + j * sg.n_chunks;
    NoteRec* base = p.notes + sg.t_off * kNotes + (int64_t)j * sg.n_rows;
    double* base_on = p.onsets + sg.t_off * kNotes + (int64_t)j * sg.n_rows;
    int64_t total = 0;
    for (int c = 0; c < sg.n_chunks; ++c) {
        const int cnt = p.chunk_count[ce + c];
        const int64_t src = (int64_t)c * kNoteChunk;
        if (cnt > 0 && src != total) {
            NoteRec r[kNoteChunk / 32];
#pragma unroll
            for (int k = 0; k < kNoteChunk / 32; ++k)
                if (lane + 32 * k < cnt) r[k] = base[src + lane + 32 * k];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kNoteChunk / 32; ++k)
                if (lane + 32 * k < cnt) {
                    base[total + lane + 32 * k] = r[k];
                    base_on[total + lane + 32 * k] = r[k].onset;
                }
            __syncwarp();
        }
        if (cnt > 0 && total > 0 && lane == 0) {   // first note of this chunk against the notes before it
            int64_t k = total;
            while (k > 0 && base_on[k - 1] > base_on[k]) {
                const NoteRec t = base[k]; base[k] = base[k - 1]; base[k - 1] = t;
                const double to = base_on[k]; base_on[k] = base_on[k - 1]; base_on[k - 1] = to;
                --k;
            }
        }
        __syncwarp();
        total += cnt;
    }
    if (lane == 0) p.counts[gw] = total;
}

// Per-song totals and their exclusive prefix (one block): out_base[s] = first sorted record of song s, song_total[s] its
// note count, song_total[n_songs] the grand total.
__global__ void __launch_bounds__(256) notes_bases_kernel(const int64_t* __restrict__ counts, int n_songs, int64_t* __restrict__ out_base,
                                                          int64_t* __restrict__ song_total) {
    __shared__ int64_t s_scan[256];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int s0 = 0; s0 < n_songs; s0 += 256) {
        const int s = s0 + threadIdx.x;
        int64_t tot = 0;
        if (s < n_songs)
            for (int j = 0; j < kNotes; ++j) tot += counts[s * kNotes + j];
        s_scan[threadIdx.x] = tot;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {   // inclusive Hillis-Steele scan
            const int64_t add = threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += add;
            __syncthreads();
        }
        if (s < n_songs) {
            out_base[s] = s_carry + s_scan[threadIdx.x] - tot;
            song_total[s] = tot;
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry += s_scan[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) song_total[n_songs] = s_carry;
}

// Rank of note k of pitch j among the song's notes: k + sum over pitches j' != j of #{notes of j' with onset < t, or onset == t
// and j' < j}.  base[j'] = first dense record of pitch j', pref[j'] = number of notes of the pitches before j'.
NOTE_HD int64_t note_rank(const double* onsets, const int64_t* base, const int64_t* pref, int j, int64_t k) {
    const double t = onsets[base[j] + k];
    int64_t rank = k;
    for (int jj = 0; jj < kNotes; ++jj) {
        if (jj == j) continue;
        const int64_t first = base[jj];
        int64_t lo = first, hi = first + (pref[jj + 1] - pref[jj]);
        if (jj < j) {  // upper bound: first onset > t
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (NOTE_LD(onsets + mid) <= t) lo = mid + 1; else hi = mid;
            }
        } else {       // lower bound: first onset >= t
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (NOTE_LD(onsets + mid) < t) lo = mid + 1; else hi = mid;
            }
        }
        rank += lo - first;
    }
    return rank;
}

// Final order of a song's notes: stable sort by onset of the pitch-major array (extractor.py:416).  Note k of pitch j
// lands at  k + sum over pitches j' != j of #{notes of j' with onset < t, or onset == t and j' < j}.
// grid (blocks per song, n_songs), grid-stride over the song's notes (the host does not know the counts at launch time).
__global__ void __launch_bounds__(256)
notes_rank_kernel(const NoteRec* __restrict__ notes, const double* __restrict__ onsets, const NotesSong* __restrict__ songs,
                  const int64_t* __restrict__ counts, const int64_t* __restrict__ out_base, NoteRec* __restrict__ sorted) {
    __shared__ int64_t s_base[kNotes];
    __shared__ int64_t s_pref[kNotes + 1];   // dense prefix of the song's per-pitch counts
    const int song = blockIdx.y;
    const NotesSong sg = songs[song];
    for (int j = threadIdx.x; j < kNotes; j += blockDim.x) s_base[j] = sg.t_off * kNotes + (int64_t)j * sg.n_rows;
    if (threadIdx.x == 0) {
        int64_t acc = 0;
        for (int j = 0; j < kNotes; ++j) { s_pref[j] = acc; acc += counts[song * kNotes + j]; }
        s_pref[kNotes] = acc;
    }
    __syncthreads();
    const int64_t n = s_pref[kNotes];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
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
        const int64_t rank = note_rank(onsets, s_base, s_pref, j, k);
        sorted[out_base[song] + rank] = notes[g];
    }
}

}  // namespace etude
