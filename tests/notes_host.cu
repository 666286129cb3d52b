// TEST INFRASTRUCTURE: the note stage's per-item logic (etude_b200/csrc/notes.cuh: notes_scan_item, notes_walk_item,
// note_rank -- __host__ __device__) compiled for the CPU, so that the chunked algorithm can be checked against the
// reference-pinned oracle without a GPU (tests/test_notes_host.py).  The kernels' orchestration is restated serially:
// transpose, scan every (pitch, chunk), walk every (pitch, chunk), compact per pitch (same rule as notes_compact_kernel),
// rank.  Built by tests/test_notes_host.py with  nvcc -x cu -O1 -Xcompiler -ffp-contract=off,-fPIC -shared.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../etude_b200/csrc/notes.cuh"

using namespace etude;

extern "C" int64_t notes_host(const float* on, const float* off, const float* mpe, const int8_t* vel, int64_t T, int note_min, double hop_sec,
                              double thr_on, double thr_off, double thr_mpe, int mode_velocity, int mode_offset, NoteRec* out) {
    const int nc = (int)((T + kNoteChunk - 1) / kNoteChunk);
    std::vector<float> t_on((size_t)T * kNotes), t_off((size_t)T * kNotes), t_mpe((size_t)T * kNotes);
    for (int64_t i = 0; i < T; ++i)
        for (int j = 0; j < kNotes; ++j) {
            t_on[(size_t)j * T + i] = on[i * kNotes + j];
            t_off[(size_t)j * T + i] = off[i * kNotes + j];
            t_mpe[(size_t)j * T + i] = mpe[i * kNotes + j];
        }
    NotesSong sg{0, T, 0, 0, nc};
    std::vector<int32_t> tab((size_t)4 * nc * kNotes, -7);
    std::vector<int64_t> counts(kNotes, 0);
    std::vector<NoteRec> slab((size_t)T * kNotes);
    std::vector<double> onsets((size_t)T * kNotes, 0.0);
    NotesParams p{};
    p.onset = t_on.data(); p.offset = t_off.data(); p.mpe = t_mpe.data(); p.velocity = vel; p.songs = &sg; p.n_songs = 1;
    p.note_min = note_min; p.hop_sec = hop_sec;
    p.thr_onset = (float)thr_on; p.thr_offset = (float)thr_off; p.thr_mpe = (float)thr_mpe;
    p.mode_velocity = mode_velocity; p.mode_offset = mode_offset;
    p.first_on = tab.data(); p.first_kept = p.first_on + (size_t)nc * kNotes; p.first_off = p.first_kept + (size_t)nc * kNotes;
    p.chunk_count = p.first_off + (size_t)nc * kNotes;
    p.counts = counts.data(); p.notes = slab.data(); p.onsets = onsets.data();
    for (int j = 0; j < kNotes; ++j)
        for (int c = 0; c < nc; ++c) notes_scan_item(p, sg, j, c);
    for (int j = kNotes - 1; j >= 0; --j)            // any order: the items are independent
        for (int c = nc - 1; c >= 0; --c) notes_walk_item(p, sg, j, c);
    // compaction per pitch (notes_compact_kernel's rule: move chunk lists together, repair the order across chunk borders)
    for (int j = 0; j < kNotes; ++j) {
        NoteRec* base = slab.data() + (size_t)j * T;
        double* base_on = onsets.data() + (size_t)j * T;
        int64_t total = 0;
        for (int c = 0; c < nc; ++c) {
            const int cnt = p.chunk_count[j * nc + c];
            const int64_t src = (int64_t)c * kNoteChunk;
            if (cnt > 0 && src != total) {
                std::vector<NoteRec> r(base + src, base + src + cnt);
                for (int k = 0; k < cnt; ++k) { base[total + k] = r[k]; base_on[total + k] = r[k].onset; }
            }
            if (cnt > 0 && total > 0) {
                int64_t k = total;
                while (k > 0 && base_on[k - 1] > base_on[k]) {
                    const NoteRec t = base[k]; base[k] = base[k - 1]; base[k - 1] = t;
                    const double to = base_on[k]; base_on[k] = base_on[k - 1]; base_on[k - 1] = to;
                    --k;
                }
            }
            total += cnt;
        }
        counts[j] = total;
    }
    std::vector<int64_t> s_base(kNotes), s_pref(kNotes + 1);
    int64_t acc = 0;
    for (int j = 0; j < kNotes; ++j) { s_base[j] = (int64_t)j * T; s_pref[j] = acc; acc += counts[j]; }
    s_pref[kNotes] = acc;
    for (int j = 0; j < kNotes; ++j)
        for (int64_t k = 0; k < counts[j]; ++k) out[note_rank(onsets.data(), s_base.data(), s_pref.data(), j, k)] = slab[s_base[j] + k];
    return acc;
}
