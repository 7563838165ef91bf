/* Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the note-event featurisation, used as the checker at
 * full BASELINE sizes and as the CPU baseline of the rasteriser in bench.py.  Never linked into the product.
 *
 * Follows, line by line, oracle/featurise.py (rasterize_sequence / rasterize_batch), i.e.
 *   tokens : EventBasedMIDIReader._parse_track  (MIDIUtil/midi_io.py:70-93), ids from MIDIUtil/defaults.py:43-58,
 *            create_timeshift_event binning int(ticks/30) (MIDIUtil/Melody.py:117-126)
 *   clock  : MelodyWriter._write_track (MIDIUtil/midi_io.py:119-127): every TIMESHIFT token adds 30*bin ticks
 *   roll   : derived piano-roll spec; parity unpinned.
 * Build: gcc -O2 -shared -fPIC -fopenmp -o oracle/_build/liboracle_raster.so oracle/raster.c  (__graft_entry__.build)
 */
#include <stdint.h>
#include <string.h>

#define MAX_TICKS 1000
#define TICKS_PER_BIN 30
#define PAD_ID 0
#define SOS_ID 1
#define NOTE_ON_FIRST 3
#define NOTE_OFF_FIRST 131
#define TIMESHIFT_FIRST 259
#define N_PITCH 128

static void one_sequence(const int32_t* dtick, const uint8_t* pitch, const uint8_t* vel, int n_ev, int res, int spq,
                         int S, int L, int velocity_roll, int32_t* tok_row, uint8_t* roll, int32_t* n_tokens) {
  int on_slice[N_PITCH];
  uint8_t on_val[N_PITCH];
  for (int p = 0; p < N_PITCH; ++p) on_slice[p] = -1;
  for (int i = 0; i <= L; ++i) tok_row[i] = i == 0 ? SOS_ID : PAD_ID;
  memset(roll, 0, (size_t)S * N_PITCH);
  long long clock = 0;
  int count = 0, open = 1;
  for (int e = 0; e < n_ev; ++e) {
    long long delta = dtick[e];
    const int p = pitch[e] & 0x7F, v = vel[e] & 0x7F;   /* 7-bit MIDI data bytes */
    long long played = 0;
    while (delta > 0) {                                   /* midi_io.py:81-83 */
      const int bin = (int)((delta % MAX_TICKS) / TICKS_PER_BIN);
      if (count < L) tok_row[1 + count] = TIMESHIFT_FIRST + bin;
      ++count;
      played += (long long)TICKS_PER_BIN * bin;           /* Melody.py:82-83 */
      delta -= MAX_TICKS;
    }
    if (count < L) tok_row[1 + count] = (v > 0 ? NOTE_ON_FIRST : NOTE_OFF_FIRST) + p;   /* midi_io.py:85-89 */
    ++count;
    clock += played;
    if (!open) continue;
    const long long s = (clock * spq) / res;
    if (s >= S) { open = 0; continue; }                   /* time-ordered: everything later is beyond the window */
    if (v > 0) {
      if (on_slice[p] >= 0) {
        long long end = s - 1 > on_slice[p] ? s - 1 : on_slice[p];
        for (long long t = on_slice[p]; t <= end; ++t) roll[t * N_PITCH + p] = on_val[p];
      }
      on_slice[p] = (int)s;
      on_val[p] = velocity_roll ? (uint8_t)v : 1;
    } else if (on_slice[p] >= 0) {
      long long end = (clock * spq + res - 1) / res - 1;  /* ceil - 1 */
      if (end < on_slice[p]) end = on_slice[p];
      if (end > S - 1) end = S - 1;
      for (long long t = on_slice[p]; t <= end; ++t) roll[t * N_PITCH + p] = on_val[p];
      on_slice[p] = -1;
    }
  }
  for (int p = 0; p < N_PITCH; ++p)
    if (on_slice[p] >= 0)
      for (int t = on_slice[p]; t < S; ++t) roll[(size_t)t * N_PITCH + p] = on_val[p];
  *n_tokens = count;
}

void oracle_rasterize(const int32_t* dtick, const uint8_t* pitch, const uint8_t* vel, const int32_t* seq_offsets,
                      int n_seq, int res, int spq, int S, int L, int velocity_roll, int32_t* tokens, uint8_t* roll,
                      int32_t* n_tokens, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads)
  for (int n = 0; n < n_seq; ++n) {
    const int a = seq_offsets[n], b = seq_offsets[n + 1];
    one_sequence(dtick + a, pitch + a, vel + a, b - a, res, spq, S, L, velocity_roll, tokens + (size_t)n * (L + 1),
                 roll + (size_t)n * S * N_PITCH, n_tokens + n);
  }
}
