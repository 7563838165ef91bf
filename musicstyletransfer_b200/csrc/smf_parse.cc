// (f3) Standard MIDI File bytes -> note-event SoA for K1.  Host-side C++ (no device work), part of libmsx.so.
//
// Replaces python-midi's `midi.read_midifile` plus the event walk of EventBasedMIDIReader.read_file / _parse_track
// (/root/reference/music_style_transfer/MIDIUtil/midi_io.py:35-68,70-93) and _extract_bpm (:16-25):
//   * every event's delta time advances the track clock (:75);
//   * only Note-On (0x9n) and Note-Off (0x8n) channel events produce an output event, with dtick = clock since the previous
//     NOTE event (:79-91), pitch = data byte 0, velocity = data byte 1 (NOTE_ON / NOTE_OFF is decided from the velocity by
//     K1, as :85-89 does);
//   * bpm comes from the first Set-Tempo meta event in (track, event) order, 120 when there is none;
//   * running status, meta events (0xFF type len data), sysex (0xF0 / 0xF7 len data) as in the SMF 1.0 specification.
// Per track the parser also returns the number of tokens A1 produces (sum over events of ceil(dtick / 1000) time shifts
// + 1 note token), so the caller applies the reference's "< 10 tokens: discard" rule (:60-63) without touching the device.
#include <stdint.h>
#include <string.h>

#include "../../include/msx.h"

void msx_set_error(const char* fmt, ...);

namespace {

struct Reader {
  const uint8_t* p;
  long long n, pos;
  bool ok;
  uint8_t u8() {
    if (pos >= n) { ok = false; return 0; }
    return p[pos++];
  }
  uint32_t be(int bytes) {
    uint32_t v = 0;
    for (int i = 0; i < bytes; ++i) v = (v << 8) | u8();
    return v;
  }
  // variable-length quantity (at most 4 bytes in a well-formed file; longer ones are consumed like python-midi does)
  long long varlen() {
    long long v = 0;
    for (int i = 0; i < 8; ++i) {
      const uint8_t b = u8();
      v = (v << 7) | (b & 0x7F);
      if (!(b & 0x80) || !ok) return v;
    }
    ok = false;
    return v;
  }
};

inline int channel_data_bytes(uint8_t status) {
  switch (status >> 4) {
    case 0x8: case 0x9: case 0xA: case 0xB: case 0xE: return 2;
    case 0xC: case 0xD: return 1;
    default: return -1;
  }
}

}  // namespace

extern "C" int msx_smf_parse(const uint8_t* bytes, long long n_bytes, long long event_capacity, int track_capacity,
                             int32_t* dtick, uint8_t* pitch, uint8_t* vel, int32_t* track_offsets, int32_t* track_tokens,
                             msx_smf_info* info) {
  if (!bytes || !info || n_bytes < 14) { msx_set_error("msx_smf_parse: not a Standard MIDI File (too short)"); return MSX_ERR_ARG; }
  if (event_capacity < 0 || track_capacity < 0 || (event_capacity > 0 && (!dtick || !pitch || !vel)) ||
      (track_capacity > 0 && (!track_offsets || !track_tokens))) {
    msx_set_error("msx_smf_parse: null output buffer");
    return MSX_ERR_ARG;
  }
  Reader r{bytes, n_bytes, 0, true};
  if (memcmp(bytes, "MThd", 4) != 0) { msx_set_error("msx_smf_parse: not a Standard MIDI File"); return MSX_ERR_ARG; }
  r.pos = 4;
  const uint32_t hlen = r.be(4);
  const uint32_t fmt = r.be(2), ntrks = r.be(2), division = r.be(2);
  if (division & 0x8000) { msx_set_error("msx_smf_parse: SMPTE time division is not supported"); return MSX_ERR_UNSUPPORTED; }
  r.pos = 8 + (long long)hlen;
  info->resolution = (int)division;
  info->format = (int)fmt;
  info->n_tracks = (int)ntrks;
  info->bpm = 120.0;                                  // MIDIUtil/defaults.py DEFAULT_BPM
  info->n_events = 0;
  bool have_tempo = false;
  long long n_ev = 0;
  if ((int)ntrks > track_capacity) {
    msx_set_error("msx_smf_parse: %u tracks exceed the caller's track capacity %d", ntrks, track_capacity);
    return MSX_ERR_ARG;
  }
  for (uint32_t t = 0; t < ntrks; ++t) {
    if (r.pos + 8 > n_bytes || memcmp(bytes + r.pos, "MTrk", 4) != 0) {
      msx_set_error("msx_smf_parse: bad track chunk at byte %lld", r.pos);
      return MSX_ERR_ARG;
    }
    r.pos += 4;
    const uint32_t tlen = r.be(4);
    const long long end = r.pos + (long long)tlen;
    if (end > n_bytes) { msx_set_error("msx_smf_parse: track %u runs past the end of the file", t); return MSX_ERR_ARG; }
    track_offsets[t] = (int32_t)n_ev;
    long long clock = 0, prev_note = 0, tokens = 0;
    int status = -1;
    while (r.pos < end) {
      clock += r.varlen();
      if (!r.ok || r.pos >= end) break;
      const uint8_t b = bytes[r.pos];
      if (b == 0xFF) {                                 // meta event: FF type len data
        r.pos += 1;
        const uint8_t mtype = r.u8();
        const long long len = r.varlen();
        if (!r.ok || r.pos + len > end) { msx_set_error("msx_smf_parse: truncated meta event in track %u", t); return MSX_ERR_ARG; }
        if (mtype == 0x51 && len >= 3 && !have_tempo) {
          const uint32_t mpqn = ((uint32_t)bytes[r.pos] << 16) | ((uint32_t)bytes[r.pos + 1] << 8) | bytes[r.pos + 2];
          if (mpqn > 0) info->bpm = 60000000.0 / (double)mpqn;
          have_tempo = true;
        }
        r.pos += len;
      } else if (b == 0xF0 || b == 0xF7) {             // sysex: F0 / F7 len data
        r.pos += 1;
        const long long len = r.varlen();
        if (!r.ok || r.pos + len > end) { msx_set_error("msx_smf_parse: truncated sysex event in track %u", t); return MSX_ERR_ARG; }
        r.pos += len;
      } else {
        if (b & 0x80) {
          status = b;
          r.pos += 1;
        } else if (status < 0) {
          msx_set_error("msx_smf_parse: running status without a status byte (track %u, byte %lld)", t, r.pos);
          return MSX_ERR_ARG;
        }
        const int nd = channel_data_bytes((uint8_t)status);
        if (nd < 0) { msx_set_error("msx_smf_parse: unsupported status byte 0x%02X in track %u", status, t); return MSX_ERR_ARG; }
        if (r.pos + nd > end) { msx_set_error("msx_smf_parse: truncated channel event in track %u", t); return MSX_ERR_ARG; }
        const uint8_t d0 = bytes[r.pos], d1 = nd > 1 ? bytes[r.pos + 1] : 0;
        if ((d0 & 0x80) || (d1 & 0x80)) {
          msx_set_error("msx_smf_parse: malformed channel event at byte %lld: data bytes must be 7-bit", r.pos);
          return MSX_ERR_ARG;
        }
        r.pos += nd;
        const int kind = status >> 4;
        if (kind == 0x9 || kind == 0x8) {
          const long long d = clock - prev_note;
          if (d > 0x7FFFFFFFll) { msx_set_error("msx_smf_parse: delta time overflows 31 bits in track %u", t); return MSX_ERR_ARG; }
          if (n_ev < event_capacity) {
            dtick[n_ev] = (int32_t)d;
            pitch[n_ev] = d0;
            vel[n_ev] = d1;
          }
          ++n_ev;
          tokens += (d > 0 ? (d + 999) / 1000 : 0) + 1;    // midi_io.py:81-89
          prev_note = clock;
        }
      }
    }
    if (!r.ok) { msx_set_error("msx_smf_parse: truncated event data in track %u", t); return MSX_ERR_ARG; }
    r.pos = end;
    track_tokens[t] = (int32_t)(tokens > 0x7FFFFFFFll ? 0x7FFFFFFF : tokens);
  }
  track_offsets[ntrks] = (int32_t)n_ev;
  info->n_events = n_ev;
  if (n_ev > event_capacity) {
    msx_set_error("msx_smf_parse: %lld note events exceed the caller's capacity %lld (info.n_events holds the need)", n_ev,
                  event_capacity);
    return MSX_ERR_ARG;
  }
  return MSX_OK;
}
