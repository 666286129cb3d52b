/* Oracle: piano-roll -> note list.  TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Plain-C restatement of AMTAPC_Extractor._mpe2note
 * (reference: etude/data/extractor.py:256-418) with the arithmetic types this
 * container's NumPy 2.3 (NEP 50) gives the reference code:
 *
 *   - rolls are np.float32, `hop_sec` is a Python float (256/16000);
 *   - threshold compares are float32 compares (Python float is "weak");
 *   - interpolated peak times (extractor.py:293/295, 323/325) are
 *       f32( f32(i*hop_sec)  -/+  f32(f32(hop_sec*0.5) * (a-b)) / (c-d) )
 *     and are widened to float64 by `float(...)`;
 *   - edge / plateau / equal-neighbour times, time_mpe and the last
 *     time_next are exact float64 products k*hop_sec.
 *
 * Peak test (extractor.py:270-286): frame i is a peak iff a[i] >= thr and,
 * scanning outward past equal values, the first different neighbour on each
 * side is smaller (or the array ends).  Restated per maximal run of equal
 * values, which is the same predicate in O(T).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t pitch;
    int32_t velocity;
    double onset;
    double offset;
} oracle_note_t;

typedef struct {
    int64_t loc;
    double time;
} peak_t;

static int64_t find_peaks(const float *a, int64_t T, int64_t stride, float thr, double hop_sec, peak_t *out) {
    int64_t n = 0;
    const float half_hop = (float)(hop_sec * 0.5);
    int64_t s = 0;
    while (s < T) {
        float v = a[s * stride];
        int64_t e = s;
        while (e + 1 < T && a[(e + 1) * stride] == v) e++;
        /* run [s, e] of equal values */
        if (v >= thr) {
            int left = (s == 0) || (v > a[(s - 1) * stride]);
            int right = (e == T - 1) || (v > a[(e + 1) * stride]);
            if (left && right) {
                for (int64_t i = s; i <= e; i++) {
                    double t;
                    if (i == 0 || i == T - 1) {
                        t = (double)i * hop_sec;
                    } else {
                        float p = a[(i - 1) * stride], q = a[(i + 1) * stride];
                        if (p == q) {
                            t = (double)i * hop_sec;
                        } else if (p > q) {
                            float num = half_hop * (p - q);
                            float frac = num / (v - q);
                            float tf = (float)((double)i * hop_sec) - frac;
                            t = (double)tf;
                        } else {
                            float num = half_hop * (q - p);
                            float frac = num / (v - p);
                            float tf = (float)((double)i * hop_sec) + frac;
                            t = (double)tf;
                        }
                    }
                    out[n].loc = i;
                    out[n].time = t;
                    n++;
                }
            }
        }
        s = e + 1;
    }
    return n;
}

static int cmp_onset_stable(const void *x, const void *y) {
    const oracle_note_t *a = *(const oracle_note_t *const *)x, *b = *(const oracle_note_t *const *)y;
    if (a->onset < b->onset) return -1;
    if (a->onset > b->onset) return 1;
    return (a < b) ? -1 : (a > b);   /* original (pitch-major) order breaks ties */
}

/* mode_velocity: 0 = 'ignore_zero', 1 = 'org'.  mode_offset: 0 = 'shorter', 1 = 'longer', 2 = 'offset'.
 * Rolls are [T, num_note] row-major.  Returns the number of notes written to *out (malloc'd, caller frees
 * with oracle_free), sorted like extractor.py:416 (stable by pitch, then stable by onset). */
int64_t oracle_mpe2note(const float *onset, const float *offset, const float *mpe, const int8_t *velocity,
                        int64_t T, int num_note, int note_min, double hop_sec, double thred_onset,
                        double thred_offset, double thred_mpe, int mode_velocity, int mode_offset,
                        oracle_note_t **out) {
    const float thr_on = (float)thred_onset, thr_off = (float)thred_offset, thr_mpe = (float)thred_mpe;
    int64_t cap = 1024, n = 0;
    oracle_note_t *notes = (oracle_note_t *)malloc(cap * sizeof(*notes));
    peak_t *pon = (peak_t *)malloc((T > 0 ? T : 1) * sizeof(peak_t));
    peak_t *poff = (peak_t *)malloc((T > 0 ? T : 1) * sizeof(peak_t));
    for (int j = 0; j < num_note; j++) {
        int64_t n_on = find_peaks(onset + j, T, num_note, thr_on, hop_sec, pon);
        int64_t n_off = find_peaks(offset + j, T, num_note, thr_off, hop_sec, poff);
        double time_next = 0.0, time_offset = 0.0, time_mpe = 0.0;
        int64_t idx_off = 0;
        for (int64_t k = 0; k < n_on; k++) {
            int64_t loc_onset = pon[k].loc, loc_next;
            double time_onset = pon[k].time;
            if (k + 1 < n_on) {
                loc_next = pon[k + 1].loc;
                time_next = pon[k + 1].time;
            } else {
                loc_next = T;
                time_next = (double)(loc_next - 1) * hop_sec;
            }
            int64_t loc_offset = loc_onset + 1;
            int flag_offset = 0;
            while (idx_off < n_off && poff[idx_off].loc <= loc_onset) idx_off++;   /* first offset peak after onset */
            if (idx_off < n_off) {
                loc_offset = poff[idx_off].loc;
                time_offset = poff[idx_off].time;
                flag_offset = 1;
            }
            if (loc_offset > loc_next) {
                loc_offset = loc_next;
                time_offset = time_next;
            }
            int64_t loc_mpe = loc_onset + 1;
            int flag_mpe = 0;
            for (int64_t ii = loc_onset + 1; ii < loc_next; ii++) {
                if (mpe[ii * num_note + j] < thr_mpe) {
                    loc_mpe = ii;
                    flag_mpe = 1;
                    time_mpe = (double)loc_mpe * hop_sec;
                    break;
                }
            }
            int velocity_value = (int)velocity[loc_onset * num_note + j];
            double offset_value;
            if (!flag_offset && !flag_mpe) offset_value = time_next;
            else if (flag_offset && !flag_mpe) offset_value = time_offset;
            else if (!flag_offset && flag_mpe) offset_value = time_mpe;
            else if (mode_offset == 2) offset_value = time_offset;
            else if (mode_offset == 1) offset_value = (loc_offset >= loc_mpe) ? time_offset : time_mpe;
            else offset_value = (loc_offset <= loc_mpe) ? time_offset : time_mpe;
            if (mode_velocity != 0 || velocity_value > 0) {
                if (n == cap) {
                    cap *= 2;
                    notes = (oracle_note_t *)realloc(notes, cap * sizeof(*notes));
                }
                notes[n].pitch = j + note_min;
                notes[n].velocity = velocity_value;
                notes[n].onset = time_onset;
                notes[n].offset = offset_value;
                n++;
            }
            if (n > 1 && notes[n - 1].pitch == notes[n - 2].pitch && notes[n - 1].onset < notes[n - 2].offset)
                notes[n - 2].offset = notes[n - 1].onset;
        }
    }
    free(pon);
    free(poff);
    /* sorted(sorted(a, key=pitch), key=onset): notes are already pitch-major; stable sort by onset. */
    oracle_note_t **idx = (oracle_note_t **)malloc((n > 0 ? n : 1) * sizeof(*idx));
    for (int64_t i = 0; i < n; i++) idx[i] = &notes[i];
    qsort(idx, (size_t)n, sizeof(*idx), cmp_onset_stable);
    oracle_note_t *sorted = (oracle_note_t *)malloc((n > 0 ? n : 1) * sizeof(*sorted));
    for (int64_t i = 0; i < n; i++) sorted[i] = *idx[i];
    free(idx);
    free(notes);
    *out = sorted;
    return n;
}

void oracle_free(void *p) { free(p); }
