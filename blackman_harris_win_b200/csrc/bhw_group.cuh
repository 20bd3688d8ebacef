// bhw_group.cuh - per-thread bodies of the group synthesis kernel (k_synth_group) and of the
// half-period table pyramid it reads.  Same conventions as bhw_device.cuh: __device__ under nvcc,
// plain inline functions under g++ for tests/hostcheck (test infrastructure; the product has no CPU path).
//
// A *family* is one sin/cos source at one DAT_WIDTH (cordic_dds or the HLS cordic; src/cordic_dds.vhd:97-249,
// hls/windows/win_function.cpp:47-156).  Such a source only sees left-aligned phase bits
// (src/cordic_dds.vhd:159-166), so its cosine sequence at PHASE_WIDTH L-1 is every second value of the
// sequence at PHASE_WIDTH L, and the second half of a period is the exact negation of the first
// (source_antisymmetric()).  All windows of a family - whatever their PHI_WIDTH, entity or ports - therefore
// read ONE table, the half-period pyramid
//     H[2^(L-1) + i] = cos(phase i of L bits) << tshift,   0 <= i < 2^(L-1),   lmin <= L <= top,
// (a binary heap: level L starts at word 2^(L-1); the whole pyramid is 2^top words), built once per step by
// one evaluation per quarter-wave phase of the top level.  Harmonic k = 2^a * b (b odd) of a 2^pw-point window
// reads level min(top, pw - a) at stride b: even harmonics become contiguous reads of a smaller level.
//
// A *group* is every window of a family with the same number of terms and tail: one launch of
// k_synth_group generates all of them (the win_selector sweep, src/win_selector.vhd:93-199: 10 variants x
// PHI_WIDTH 4..26 = 5 launches instead of 100), with the table staged in shared memory once per CTA:
//   G_HALF32 : the top level's half period as int32 (<= 2^15 words ... 192 KB)
//   G_Q16    : the top level's quarter wave as two uint16 arrays (cos, sin; value + bias), for DAT_WIDTH <= 17
//              where |value| <= 2^15 + eps: a 2^17-entry table (512 KB as int32) in 128 KB
//   G_GLOBAL : nothing staged, gathers from the pyramid through L1/L2
#pragma once
#include <stdint.h>

#include "bhw_device.cuh"

namespace bhw {

enum : int { G_HALF32 = 0, G_Q16 = 1, G_GLOBAL = 2 };

struct GroupShape {
  uint32_t m;
  uint32_t top;                  // top level of the pyramid = log2(entries of a full period at full resolution)
  uint32_t lmin;                 // lowest level present
  uint32_t rc, rcn;              // product rounding addends (WinRec comment): b(P), and -b(-P)
  int32_t lsh, rsh;              // final shifts
  uint32_t interleave;           // 1: tiles interleaved over the grid, 0: one contiguous share per CTA
  uint32_t tmul, tbias;          // G_Q16: table entry = stored * tmul - tbias  (tmul = 2^tshift, tbias = bias << tshift)
  const int32_t* pyr;            // the pyramid (global memory)
  const uint16_t* q16;           // G_Q16: cos16[Q] then sin16[Q], Q = 2^(top-2)
};

// One window (or tile range of a window) of a group launch.  Units are tiles of 256 samples - of 256 sample
// pairs (n, n + N/2) in a paired launch.  wins[nwin] is a sentinel carrying the total unit count.
struct GroupWin {
  uint32_t unit_begin;           // first unit of this window in the launch
  uint32_t pw;                   // PHI_WIDTH
  uint32_t rec;                  // window record (ports)
  uint32_t tile_first;           // first tile of the window this launch generates (0 for whole windows)
  int64_t out_off;               // element offset, relative to the launch's out pointer, of the window's sample 0
  uint64_t pad;
};

// two's-complement trailing zeros of the harmonic number (k = 1..6)
BHW_HD constexpr uint32_t harmonic_log2(int k) { return (k & 1) ? 0u : (k & 2) ? 1u : 2u; }

// ---- synthesis -----------------------------------------------------------------------------------
// One lane's share of a tile: samples nbase + lane + 32*j (j = 0..7) -> va[j], and their partners half a
// window later -> vb[j] when PAIR.  `tab`: the staged image (G_HALF32: int32 half period of level `top`;
// G_Q16: the two uint16 quarter waves) or the pyramid (G_GLOBAL).
// Harmonic k's 32-bit phase is n * (k << (32 - pw)); a tile in which that phase stays inside one half
// period (one quadrant for G_Q16) has a warp-uniform sign (and table), folded into the coefficient.
template <int M, int TAB, bool PAIR>
BHW_HD void group_lane_tile(const GroupShape& sh, uint32_t pw, const int32_t* A, int32_t S0, const void* tab,
                            uint32_t nbase, uint32_t lane, int32_t* va, int32_t* vb) {
  uint32_t Sa[kBankJ], Sb[kBankJ];
#pragma unroll
  for (int j = 0; j < kBankJ; ++j) { Sa[j] = (uint32_t)S0; Sb[j] = (uint32_t)S0; }
  const uint32_t sl = 32u - pw;
  const uint32_t n = nbase + lane;
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t ks = (uint32_t)k << sl;
    const uint32_t ph_first = nbase * ks;
    const uint32_t span = (uint32_t)(kBankTile - 1) * ks;                 // meaningful when it does not overflow
    if (TAB == G_Q16) {
      // image: cos16[Q] then sin16[Q], Q = 2^(top-2).  Quadrant q = phase >> 30 reads sin when q is odd, negated
      // when q is 1 or 2; element index = (q & 1) * Q + (quarter-wave phase >> (32 - top)) = (phase & 0x7FFFFFFF)
      // >> (32 - top): the table select rides in phase bit 30.
      const uint16_t* T16 = reinterpret_cast<const uint16_t*>(tab);
      const uint32_t rsh = 32u - sh.top;
      const bool uni = ((((uint32_t)k * (uint32_t)(kBankTile - 1)) >> (pw - 2)) == 0u) &&
                       ((ph_first & 0x3FFFFFFFu) + span < 0x40000000u);
      if (uni) {
        const uint32_t q = ph_first >> 30;
        const int32_t Ak = ((q + 1u) & 2u) ? -A[k] : A[k];
        uint32_t phx = (n * ks) & 0x7FFFFFFFu;                   // stays inside the quadrant: no carry into bit 30
        const uint32_t st = ks << 5;
#pragma unroll
        for (int j = 0; j < kBankJ; ++j) {
          BHW_CHECK((phx >> rsh) < (1u << (sh.top - 1)));
          const int32_t c2 = (int32_t)((uint32_t)T16[phx >> rsh] * sh.tmul - sh.tbias);
          const int64_t P = (int64_t)Ak * (int64_t)c2;
          const uint32_t ba = (uint32_t)((P + (int64_t)(uint64_t)sh.rc) >> 32);
          Sa[j] = (k & 1) ? Sa[j] - ba : Sa[j] + ba;
          if (PAIR) Sb[j] += (k & 1) ? (uint32_t)((P + (int64_t)(uint64_t)sh.rcn) >> 32) : ba;
          phx += st;
        }
      } else {
        uint32_t ph = n * ks;
        const uint32_t st = ks << 5;
#pragma unroll
        for (int j = 0; j < kBankJ; ++j) {
          const uint32_t q = ph >> 30;
          BHW_CHECK(((ph & 0x7FFFFFFFu) >> rsh) < (1u << (sh.top - 1)));
          const uint32_t u = T16[(ph & 0x7FFFFFFFu) >> rsh];
          int32_t c2 = (int32_t)(u * sh.tmul - sh.tbias);
          c2 = ((q + 1u) & 2u) ? -c2 : c2;
          const int64_t P = (int64_t)A[k] * (int64_t)c2;
          const uint32_t ba = (uint32_t)((P + (int64_t)(uint64_t)sh.rc) >> 32);
          Sa[j] = (k & 1) ? Sa[j] - ba : Sa[j] + ba;
          if (PAIR) Sb[j] += (k & 1) ? (uint32_t)((P + (int64_t)(uint64_t)sh.rcn) >> 32) : ba;
          ph += st;
        }
      }
    } else {
      // Level this harmonic reads (G_HALF32: the staged top level for every harmonic).  Element index inside
      // the pyramid (heap layout: level L starts at word 2^(L-1)): 2^(L-1) + (half-period phase >> (31 - L))
      // = (phase | 2^31) >> (32 - L) - the level's offset rides in the top phase bit.
      uint32_t L = sh.top;
      const int32_t* T = reinterpret_cast<const int32_t*>(tab);
      bool lin = false;
      if (TAB == G_GLOBAL) {
        const uint32_t want = pw - harmonic_log2(k);
        lin = want <= sh.top;                                    // the phase step is a whole number of entries
        L = lin ? want : sh.top;
      } else {
        T -= (1u << (L - 1));                                    // the staged image is level `top` alone
      }
      const uint32_t rsh = 32u - L;
      const bool uni = ((((uint32_t)k * (uint32_t)(kBankTile - 1)) >> (pw - 1)) == 0u) &&
                       ((ph_first & 0x7FFFFFFFu) + span < 0x80000000u);
      if (uni) {
        const int32_t Ak = (ph_first >> 31) ? -A[k] : A[k];
        uint32_t phx = (n * ks) | 0x80000000u;                   // stays inside the half period: no carry into bit 31
        if (TAB == G_GLOBAL && lin) {
          // whole-entry steps: sample lane + 32*j reads entry (first + b*(lane + 32*j)), b = k >> log2: one address,
          // compile-time offsets
          const int32_t* Tl = T + (phx >> rsh);
          const uint32_t b = (uint32_t)k >> harmonic_log2(k);   // a constant once the harmonic loop is unrolled
#pragma unroll
          for (int j = 0; j < kBankJ; ++j) {
            BHW_CHECK((phx >> rsh) + (uint32_t)(32 * j) * b >= (1u << (L - 1)) && (phx >> rsh) + (uint32_t)(32 * j) * b < (1u << L));
            const int32_t c2 = Tl[(uint32_t)(32 * j) * b];
            const int64_t P = (int64_t)Ak * (int64_t)c2;
            const uint32_t ba = (uint32_t)((P + (int64_t)(uint64_t)sh.rc) >> 32);
            Sa[j] = (k & 1) ? Sa[j] - ba : Sa[j] + ba;
            if (PAIR) Sb[j] += (k & 1) ? (uint32_t)((P + (int64_t)(uint64_t)sh.rcn) >> 32) : ba;
          }
        } else {
          const uint32_t st = ks << 5;
#pragma unroll
          for (int j = 0; j < kBankJ; ++j) {
            BHW_CHECK((phx >> rsh) >= (1u << (L - 1)) && (phx >> rsh) < (1u << L) && L >= sh.lmin);
            const int32_t c2 = T[phx >> rsh];
            const int64_t P = (int64_t)Ak * (int64_t)c2;
            const uint32_t ba = (uint32_t)((P + (int64_t)(uint64_t)sh.rc) >> 32);
            Sa[j] = (k & 1) ? Sa[j] - ba : Sa[j] + ba;
            if (PAIR) Sb[j] += (k & 1) ? (uint32_t)((P + (int64_t)(uint64_t)sh.rcn) >> 32) : ba;
            phx += st;
          }
        }
      } else {
        uint32_t ph = n * ks;
        const uint32_t st = ks << 5;
#pragma unroll
        for (int j = 0; j < kBankJ; ++j) {
          BHW_CHECK(((ph | 0x80000000u) >> rsh) >= (1u << (L - 1)) && ((ph | 0x80000000u) >> rsh) < (1u << L) && L >= sh.lmin);
          const int32_t t = T[(ph | 0x80000000u) >> rsh];
          const int32_t c2 = (ph >> 31) ? -t : t;
          const int64_t P = (int64_t)A[k] * (int64_t)c2;
          const uint32_t ba = (uint32_t)((P + (int64_t)(uint64_t)sh.rc) >> 32);
          Sa[j] = (k & 1) ? Sa[j] - ba : Sa[j] + ba;
          if (PAIR) Sb[j] += (k & 1) ? (uint32_t)((P + (int64_t)(uint64_t)sh.rcn) >> 32) : ba;
          ph += st;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kBankJ; ++j) {
    va[j] = (int32_t)(Sa[j] << sh.lsh) >> sh.rsh;
    if (PAIR) vb[j] = (int32_t)(Sb[j] << sh.lsh) >> sh.rsh;
  }
}

// window of unit u: last w with wins[w].unit_begin <= u (wins[nwin] is the sentinel)
BHW_HD uint32_t group_find_window(const GroupWin* wins, uint32_t nwin, uint32_t u) {
  uint32_t lo = 0, hi = nwin - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1) >> 1;
    if (wins[mid].unit_begin <= u) lo = mid; else hi = mid - 1;
  }
  return lo;
}

}  // namespace bhw
