// ptap_tpl_host.h — host side of the TEMPLATE numeric PtAP (ptap_tpl.cuh): compiles the product terms of one
// output row of A_b = R A_f M (R = M^T, reference la_utils.py:165-182) into a gather program, and interprets such a
// program on the CPU (the interpreter is what the CPU tests and iife_tpl_emulate_row use; the CUDA kernel in
// ptap_tpl.cuh executes the same program, step for step, with one lane per program column).
//
// Why: on meshes with any regularity (the XTK foreground of the reference is a refinement of a structured
// background grid) almost all output rows have the SAME relative structure — same operand row lengths, same
// destination slots of every product term, same values of M — and differ only in WHERE their operand rows of A_f
// start.  Such rows share one template; everything that does not depend on A_f's values is precomputed once per
// template on the host, where there is time to pack it well:
//
//   staging     S[slot_p] = w[q_p] * A_f.val[beg[q_p] + e_p]      p over the T1 stage-1 terms (coalesced by operand row)
//   stage 1     O1[f] = sum of S[src] over a lane's run of terms    (the intermediate row (R A_f)[i,:], never in HBM)
//   stage 2     O2[f] = sum of coef * O1[src]                       (coef = the M value of the term)
//   write       A_b.val[row i] = O2[1..n2]
//
// Stages 1 and 2 are GATHERS: the terms of one destination are consecutive in one lane's program, accumulate in a
// register and are stored once ("flush"); a destination with more terms than the step count S is split into
// pieces that flush into extra slots and are added in a fixed order afterwards.  No shared-memory read-modify-write,
// no privatised accumulator copies, no hashing, no column indices, no atomics; the sum order is fixed by the
// program, so results are bit-reproducible.
//
// Shared-memory banks.  A step reads 32 fp64 words; the hardware serves it in two phases of 16 lanes, each conflict
// free iff its words lie in 16 distinct 8-byte banks (index mod 16).
//   * Stage 1 reads every staged value exactly once, and the staging writes it exactly once, 16 consecutive terms
//     per phase.  A term is therefore an EDGE between its write phase and its read phase; both have at most 16
//     members, so by Koenig's theorem the edges of this bipartite multigraph can be coloured with 16 colours such
//     that no phase sees a colour twice: colour = bank.  edge_colour_16 computes it (alternating-path recolouring) and
//     the staged value of term p lives at S[16 * rank + colour]: stage-1 reads AND staging writes are conflict
//     free for any program.
//   * Stage 2 reads intermediate entries several times each (once per entry of the M row), so no such guarantee
//     exists; the order of the terms inside a run and the order of a lane's runs are free, and a maximum bipartite
//     matching (lanes x banks) per phase picks terms in distinct banks where it can, on a layout of the intermediate
//     row that a few rounds of local search have adapted to the schedule.
//
// Plain C++ (no CUDA): included by ptap.cu.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace iife {
namespace tpl {

constexpr int MAX_N0 = 64;     // operand rows of stage 1 (entries of R[i,:])
constexpr int MAX_T1 = 1023;   // stage-1 product terms
constexpr int MAX_N1 = 256;    // intermediate row entries
constexpr int MAX_N2 = 256;    // output row entries
constexpr int MAX_T2 = 4095;   // stage-2 product terms
constexpr int MAX_OUT = 1023;  // entries of an out buffer (byte offsets must fit 16 bits)
constexpr uint32_t NO_FLUSH = 0xFFFFu;
constexpr uint32_t STG_PAD = 0xFFFFFFFFu;

// what the device extracts for the representative row of a template
struct Raw {
  int n0 = 0, n1 = 0, n2 = 0;
  std::vector<int> len1;       // [n0]  length of A_f row j_q
  std::vector<double> w;       // [n0]  R[i, j_q]
  std::vector<uint8_t> slot1;  // [T1]  rank of the term's column in the sorted intermediate row
  std::vector<int> len2;       // [n1]  length of M row k_q
  std::vector<double> mval;    // [T2]  M[k_q, e]
  std::vector<uint8_t> slot2;  // [T2]  rank of the term's column in the sorted output row
};

// 16-byte aligned sections inside one blob; all offsets in bytes from the start of the blob (= this header)
struct Header {
  int n0, T1, stg_steps, n1, n2, S1, S2;
  int ng1, nx1, ng2, nx2;  // ng = rounds of the combine pass, nx = extra slots
  int off_stg, off_w, off_p1, off_g1, off_c2, off_p2, off_g2;
  int blob_bytes;
  int ext1, ext2;  // index of the first extra slot in O1 / O2 (the destinations live below it)
  int wc2;         // 1: every split destination of stage 2 has at most 2 extra pieces and off_xw holds, per output
  int off_xw;      //    entry, the O2 indices of its extras (two 16-bit fields; 0 = the zero slot O2[0])
  int pad;
};
static_assert(sizeof(Header) == 96, "Header layout");

struct Term {
  int dest;     // destination 0..n_dest-1
  int src;      // LOGICAL source 0..n_src-1 (staged term p in stage 1, intermediate entry q in stage 2)
  double coef;  // stage 2 only
};

// a schedule: which term every lane processes at every step, and what it flushes afterwards
struct Schedule {
  int S = 0, n_extra = 0;
  std::vector<int> term;      // [S*32]  term index or -1 (padding)
  std::vector<int> flush;     // [S*32]  logical out id (destination d, or n_dest + extra) or -1
  std::vector<int> gd, gptr;  // split destinations: dest gd[g] += extras gptr[g] .. gptr[g+1]
  double lane_use = 0.0;      // terms / (32 S)
};

struct Piece {
  int dest, first, count, index;  // terms by_dest[dest][first .. first+count), index-th piece of dest
};

// Bin-pack the destinations' term runs into 32 lane programs of S steps, S as small as possible, then order the
// terms.  bank_of == nullptr: runs and terms in their given order (ascending operand order: the order the reference's
// row-wise Gustavson product adds them in).  Otherwise bank_of[src] is the shared-memory bank of every logical source
// and a maximum matching per phase avoids bank conflicts where it can.  Returns false if a destination has no term.
inline bool build_schedule(int n_dest, const std::vector<Term> &terms, const int *bank_of, Schedule &out, bool avoid_splits = false) {
  std::vector<std::vector<int>> by_dest((size_t)n_dest);
  for (size_t t = 0; t < terms.size(); ++t) {
    if (terms[t].dest < 0 || terms[t].dest >= n_dest) return false;
    by_dest[(size_t)terms[t].dest].push_back((int)t);
  }
  for (int d = 0; d < n_dest; ++d)
    if (by_dest[(size_t)d].empty()) return false;
  const int N = (int)terms.size();
  std::vector<Piece> pieces;
  std::vector<int> lane_of;  // per piece
  int S = std::max(1, (N + 31) / 32);
  if (avoid_splits) {  // a destination is only split when it has more than twice the mean lane load
    int mx = 0;
    for (int d = 0; d < n_dest; ++d) mx = std::max(mx, (int)by_dest[(size_t)d].size());
    if (mx <= 2 * S) S = std::max(S, mx);
  }
  for (;; ++S) {
    pieces.clear();
    for (int d = 0; d < n_dest; ++d) {
      const int c = (int)by_dest[(size_t)d].size();
      int idx = 0;
      for (int f = 0; f < c; f += S) pieces.push_back({d, f, std::min(S, c - f), idx++});
    }
    std::vector<int> order(pieces.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return pieces[(size_t)x].count > pieces[(size_t)y].count; });
    int room[32];
    for (int l = 0; l < 32; ++l) room[l] = S;
    lane_of.assign(pieces.size(), -1);
    bool ok = true;
    for (int k : order) {
      int best = -1;
      for (int l = 0; l < 32; ++l)  // best fit: the fullest lane that still takes the piece
        if (room[l] >= pieces[(size_t)k].count && (best < 0 || room[l] < room[best])) best = l;
      if (best < 0) {
        ok = false;
        break;
      }
      room[best] -= pieces[(size_t)k].count;
      lane_of[(size_t)k] = best;
    }
    if (ok) break;
  }
  // extras: every piece after the first of its destination flushes into its own slot
  int n_extra = 0;
  std::vector<int> extra_id(pieces.size(), -1);
  out.gd.clear();
  out.gptr.clear();
  out.gptr.push_back(0);
  {
    size_t k = 0;
    while (k < pieces.size()) {
      size_t k2 = k;
      while (k2 < pieces.size() && pieces[k2].dest == pieces[k].dest) ++k2;
      if (k2 - k > 1) {
        for (size_t p = k + 1; p < k2; ++p) extra_id[p] = n_extra++;
        out.gd.push_back(pieces[k].dest);
        out.gptr.push_back(n_extra);
      }
      k = k2;
    }
  }
  out.S = S;
  out.n_extra = n_extra;
  out.term.assign((size_t)S * 32, -1);
  out.flush.assign((size_t)S * 32, -1);
  out.lane_use = (double)N / (32.0 * S);
  auto flush_id = [&](int pk) { return pieces[(size_t)pk].index == 0 ? pieces[(size_t)pk].dest : n_dest + extra_id[(size_t)pk]; };
  std::vector<std::vector<int>> lane_pieces(32);
  for (size_t k = 0; k < pieces.size(); ++k) lane_pieces[(size_t)lane_of[k]].push_back((int)k);
  if (!bank_of) {  // given order
    for (int l = 0; l < 32; ++l) {
      int s = 0;
      for (int pk : lane_pieces[(size_t)l]) {
        const Piece &pc = pieces[(size_t)pk];
        for (int e = 0; e < pc.count; ++e, ++s) {
          out.term[(size_t)s * 32 + l] = by_dest[(size_t)pc.dest][(size_t)(pc.first + e)];
          if (e == pc.count - 1) out.flush[(size_t)s * 32 + l] = flush_id(pk);
        }
      }
    }
    return true;
  }
  // ---- bank-aware order: per phase (16 lanes of one step) a maximum matching between lanes and banks
  std::vector<std::vector<int>> left(pieces.size());  // remaining term indices of every piece
  for (size_t k = 0; k < pieces.size(); ++k)
    for (int e = 0; e < pieces[k].count; ++e) left[k].push_back(by_dest[(size_t)pieces[k].dest][(size_t)(pieces[k].first + e)]);
  int cur[32];
  for (int l = 0; l < 32; ++l) cur[l] = -1;
  for (int s = 0; s < S; ++s) {
    for (int half = 0; half < 2; ++half) {
      int load[16] = {0};
      int order[16], n_opt[16];
      bool idle_any = false;
      for (int k = 0; k < 16; ++k) {
        const int l = half * 16 + k;
        int opts = 0;
        if (cur[l] >= 0) opts = (int)left[(size_t)cur[l]].size();
        else
          for (int pk : lane_pieces[(size_t)l]) opts += (int)left[(size_t)pk].size();
        n_opt[k] = opts;
        if (opts == 0) idle_any = true;
      }
      // candidates of every lane, one per bank (the first term found in that bank): a bipartite graph lanes x banks
      int cand_piece[16][16], cand_pos[16][16];
      for (int k = 0; k < 16; ++k) {
        for (int b = 0; b < 16; ++b) cand_piece[k][b] = -1;
        const int l = half * 16 + k;
        if (n_opt[k] == 0) continue;
        auto scan = [&](int pk) {
          const std::vector<int> &rem = left[(size_t)pk];
          for (size_t z = 0; z < rem.size(); ++z) {
            const int b = bank_of[terms[(size_t)rem[z]].src] & 15;
            if (cand_piece[k][b] < 0) {
              cand_piece[k][b] = pk;
              cand_pos[k][b] = (int)z;
            }
          }
        };
        if (cur[l] >= 0) scan(cur[l]);
        else
          for (int pk : lane_pieces[(size_t)l])
            if (!left[(size_t)pk].empty()) scan(pk);
      }
      int bank_owner[16], lane_bank[16];
      for (int b = 0; b < 16; ++b) bank_owner[b] = -1;
      for (int k = 0; k < 16; ++k) lane_bank[k] = -1;
      (void)idle_any;  // idle lanes read a dummy word in a bank the phase leaves free
      {
        int n_banks[16];
        for (int k = 0; k < 16; ++k) {
          n_banks[k] = 0;
          for (int b = 0; b < 16; ++b) n_banks[k] += cand_piece[k][b] >= 0;
          order[k] = k;
        }
        std::stable_sort(order, order + 16, [&](int x, int y) { return n_banks[x] < n_banks[y]; });
        struct Aug {
          static bool go(int k, int (*cp)[16], int *owner, int *lb, bool *vis) {
            for (int b = 0; b < 16; ++b) {
              if (cp[k][b] < 0 || vis[b]) continue;
              vis[b] = true;
              if (owner[b] == 16) continue;
              if (owner[b] < 0 || go(owner[b], cp, owner, lb, vis)) {
                owner[b] = k;
                lb[k] = b;
                return true;
              }
            }
            return false;
          }
        };
        for (int oi = 0; oi < 16; ++oi) {
          const int k0 = order[oi];
          if (n_opt[k0] == 0) continue;
          bool visited[16] = {false};
          Aug::go(k0, cand_piece, bank_owner, lane_bank, visited);
        }
      }
      for (int b = 0; b < 16; ++b)
        if (bank_owner[b] >= 0) load[b] = 1;
      for (int k = 0; k < 16; ++k) {
        const int l = half * 16 + k;
        if (n_opt[k] == 0) continue;
        int b = lane_bank[k];
        if (b < 0) {  // unmatched: the least loaded bank among its candidates
          for (int bb = 0; bb < 16; ++bb)
            if (cand_piece[k][bb] >= 0 && (b < 0 || load[bb] < load[b])) b = bb;
          load[b]++;
        }
        const int pk = cand_piece[k][b], pos = cand_pos[k][b];
        std::vector<int> &rem = left[(size_t)pk];
        out.term[(size_t)s * 32 + l] = rem[(size_t)pos];
        rem.erase(rem.begin() + pos);
        if (rem.empty()) {
          out.flush[(size_t)s * 32 + l] = flush_id(pk);
          cur[l] = -1;
        } else {
          cur[l] = pk;
        }
      }
    }
  }
  return true;
}

// Proper edge colouring with 16 colours of a bipartite multigraph of maximum degree <= 16 (Koenig): edge e joins
// left vertex lv[e] and right vertex rv[e]; colour[e] in 0..15, no vertex sees a colour twice.
inline bool edge_colour_16(int n_left, int n_right, const std::vector<int> &lv, const std::vector<int> &rv, std::vector<int> &colour) {
  const size_t E = lv.size();
  colour.assign(E, -1);
  std::vector<int> at_l((size_t)n_left * 16, -1), at_r((size_t)n_right * 16, -1);  // edge of each colour at each vertex
  for (size_t e = 0; e < E; ++e) {
    const int u = lv[e], v = rv[e];
    int a = -1, b = -1;
    for (int c = 0; c < 16 && a < 0; ++c)
      if (at_l[(size_t)u * 16 + c] < 0) a = c;
    for (int c = 0; c < 16 && b < 0; ++c)
      if (at_r[(size_t)v * 16 + c] < 0) b = c;
    if (a < 0 || b < 0) return false;  // degree above 16
    if (a != b) {
      // a is free at u but used at v: flip the a/b alternating path that starts at v with colour a (it cannot end at u)
      std::vector<int> path;
      int x = v, c = a;
      bool on_right = true;
      for (;;) {
        const int f = on_right ? at_r[(size_t)x * 16 + c] : at_l[(size_t)x * 16 + c];
        if (f < 0) break;
        path.push_back(f);
        x = on_right ? lv[(size_t)f] : rv[(size_t)f];
        on_right = !on_right;
        c = (c == a) ? b : a;
      }
      for (int f : path) {
        at_l[(size_t)lv[(size_t)f] * 16 + colour[(size_t)f]] = -1;
        at_r[(size_t)rv[(size_t)f] * 16 + colour[(size_t)f]] = -1;
      }
      for (int f : path) {
        colour[(size_t)f] = (colour[(size_t)f] == a) ? b : a;
        at_l[(size_t)lv[(size_t)f] * 16 + colour[(size_t)f]] = f;
        at_r[(size_t)rv[(size_t)f] * 16 + colour[(size_t)f]] = f;
      }
    }
    colour[e] = a;
    at_l[(size_t)u * 16 + a] = (int)e;
    at_r[(size_t)v * 16 + a] = (int)e;
  }
  return true;
}

struct Program {
  std::vector<unsigned char> blob;  // Header + sections
  int s_cap = 0, o1_cap = 0, o2_cap = 0;  // entries of S / O1 / O2 this template needs
  int n0 = 0;                              // operand rows of stage 1
  double use1 = 0.0, use2 = 0.0;
  double conf1 = 0.0, conf2 = 0.0;  // mean conflict degree of the gather reads of the two stages (1.0 = none)
};

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// phases needed by the reads of one stage given the buffer index of every logical source (idle lanes read a dummy
// word in a free bank)
inline double stage_conflicts(const Schedule &sc, const std::vector<Term> &terms, const std::vector<int> &index_of) {
  long long wf = 0;
  for (int s = 0; s < sc.S; ++s)
    for (int half = 0; half < 2; ++half) {
      int word[16], cnt[16] = {0}, mx = 1;
      for (int k = 0; k < 16; ++k) {
        const int t = sc.term[(size_t)s * 32 + half * 16 + k];
        word[k] = t < 0 ? -1 : index_of[(size_t)terms[(size_t)t].src];
        if (t < 0) continue;
        bool dup = false;
        for (int b = 0; b < k; ++b) dup = dup || word[b] == word[k];
        if (!dup) mx = std::max(mx, ++cnt[word[k] & 15]);
      }
      wf += mx;
    }
  return (double)wf / (2.0 * std::max(sc.S, 1));
}

// false: the row does not fit the template kernel's limits (it stays on the per-row kernels)
inline bool compile(const Raw &r, Program &out) {
  if (r.n0 < 1 || r.n0 > MAX_N0 || r.n1 < 1 || r.n1 > MAX_N1 || r.n2 < 1 || r.n2 > MAX_N2) return false;
  int T1 = 0, T2 = 0;
  for (int q = 0; q < r.n0; ++q) {
    if (r.len1[(size_t)q] < 0 || r.len1[(size_t)q] > 255) return false;
    T1 += r.len1[(size_t)q];
  }
  for (int q = 0; q < r.n1; ++q) T2 += r.len2[(size_t)q];
  if (T1 < 1 || T1 > MAX_T1 || T2 < 1 || T2 > MAX_T2) return false;
  if ((int)r.slot1.size() != T1 || (int)r.slot2.size() != T2 || (int)r.mval.size() != T2) return false;
  const int stg_steps = (T1 + 31) / 32;
  std::vector<int> stg_q((size_t)T1), stg_e((size_t)T1);
  std::vector<Term> t1((size_t)T1), t2((size_t)T2);
  {
    int p = 0;
    for (int q = 0; q < r.n0; ++q)
      for (int e = 0; e < r.len1[(size_t)q]; ++e, ++p) {
        stg_q[(size_t)p] = q;
        stg_e[(size_t)p] = e;
        t1[(size_t)p] = {(int)r.slot1[(size_t)p], p, 0.0};
      }
    p = 0;
    for (int q = 0; q < r.n1; ++q)
      for (int e = 0; e < r.len2[(size_t)q]; ++e, ++p) t2[(size_t)p] = {(int)r.slot2[(size_t)p], q, r.mval[(size_t)p]};
  }
  // ---- stage 1: schedule in Gustavson order, then banks by edge colouring (write phase p / 16  x  read phase)
  Schedule s1;
  if (!build_schedule(r.n1, t1, nullptr, s1, true)) return false;  // stage 1: no combine pass in the common case
  std::vector<int> s_index((size_t)T1, 0);  // index of staged term p inside S
  {
    std::vector<int> lv((size_t)T1), rv((size_t)T1), col;
    for (int p = 0; p < T1; ++p) lv[(size_t)p] = p / 16;
    for (int s = 0; s < s1.S; ++s)
      for (int l = 0; l < 32; ++l) {
        const int t = s1.term[(size_t)s * 32 + l];
        if (t >= 0) rv[(size_t)t] = 2 * s + (l >> 4);
      }
    if (!edge_colour_16((T1 + 15) / 16, 2 * s1.S, lv, rv, col)) return false;
    int rank[16] = {0};
    for (int p = 0; p < T1; ++p) {
      s_index[(size_t)p] = 16 * rank[col[(size_t)p]] + col[(size_t)p];
      ++rank[col[(size_t)p]];
    }
  }
  int s_words = 1;  // entries of S
  for (int p = 0; p < T1; ++p) s_words = std::max(s_words, s_index[(size_t)p] + 1);
  // ---- stage 2: layout of the intermediate row adapted to a bank-aware schedule (a few rounds of local search)
  std::vector<int> o1_bank((size_t)r.n1), o1_index((size_t)r.n1);
  auto layout_from_banks = [&](const std::vector<int> &bank, std::vector<int> &index) {
    int rank[16] = {0};
    int hi = 0;
    for (int q = 0; q < r.n1; ++q) {
      const int b = bank[(size_t)q];
      index[(size_t)q] = 16 * rank[b] + b;  // (index & 15) == bank
      ++rank[b];
      hi = std::max(hi, index[(size_t)q]);
    }
    return hi + 1;  // first free index
  };
  for (int q = 0; q < r.n1; ++q) o1_bank[(size_t)q] = (q + 1) & 15;  // start from (about) the identity layout
  Schedule s2, best_s2;
  std::vector<int> best_bank = o1_bank, best_index;
  double best_conf = 1e30;
  for (int round = 0; round < 6; ++round) {
    if (!build_schedule(r.n2, t2, o1_bank.data(), s2)) return false;
    layout_from_banks(o1_bank, o1_index);
    const double conf = stage_conflicts(s2, t2, o1_index);
    if (conf < best_conf) {
      best_conf = conf;
      best_s2 = s2;
      best_bank = o1_bank;
      best_index = o1_index;
    }
    if (conf <= 1.0 + 1e-12) break;
    // local search on the banks for THIS schedule: move a source to the bank that minimises the phases it is read in
    std::vector<std::vector<int>> reads((size_t)r.n1);  // phases (2 s + half) every source is read in, ascending
    for (int s = 0; s < s2.S; ++s)
      for (int l = 0; l < 32; ++l) {
        const int t = s2.term[(size_t)s * 32 + l];
        if (t >= 0) reads[(size_t)t2[(size_t)t].src].push_back(2 * s + (l >> 4));
      }
    const int n_ph = 2 * s2.S;
    std::vector<int> cnt((size_t)n_ph * 16, 0);  // distinct sources per (phase, bank)
    auto add = [&](int q, int d) {
      int last = -1;
      for (int ph : reads[(size_t)q]) {
        if (ph == last) continue;  // several lanes of one phase reading q: one word
        cnt[(size_t)ph * 16 + o1_bank[(size_t)q]] += d;
        last = ph;
      }
    };
    for (int q = 0; q < r.n1; ++q) add(q, +1);
    // banks stay balanced (the buffer holds 16 x the fullest bank): at most ceil(n1 / 16) + 1 sources per bank
    const int bank_cap = (r.n1 + 15) / 16 + 1;
    int pop[16] = {0};
    for (int q = 0; q < r.n1; ++q) pop[o1_bank[(size_t)q]]++;
    bool moved = true;
    for (int sweep = 0; sweep < 4 && moved; ++sweep) {
      moved = false;
      for (int q = 0; q < r.n1; ++q) {
        add(q, -1);
        pop[o1_bank[(size_t)q]]--;
        int best_b = o1_bank[(size_t)q], best_cost = 1 << 30;
        for (int b = 0; b < 16; ++b) {
          if (pop[b] >= bank_cap) continue;
          int cost = 0, last = -1;
          for (int ph : reads[(size_t)q]) {
            if (ph == last) continue;
            cost += cnt[(size_t)ph * 16 + b];
            last = ph;
          }
          if (cost < best_cost || (cost == best_cost && b == o1_bank[(size_t)q])) {
            best_cost = cost;
            best_b = b;
          }
        }
        if (best_b != o1_bank[(size_t)q]) moved = true;
        o1_bank[(size_t)q] = best_b;
        pop[best_b]++;
        add(q, +1);
      }
    }
  }
  s2 = best_s2;
  o1_bank = best_bank;
  const int ext1 = layout_from_banks(o1_bank, o1_index);  // extras of stage 1 follow the destinations
  const int ext2 = 1 + r.n2;
  if (ext1 + s1.n_extra > MAX_OUT || ext2 + s2.n_extra > MAX_OUT || s_words > 8191) return false;
  // ---- emit
  auto out1_index = [&](int id) { return id < r.n1 ? o1_index[(size_t)id] : ext1 + (id - r.n1); };
  auto out2_index = [&](int id) { return id < r.n2 ? 1 + id : ext2 + (id - r.n2); };
  std::vector<uint32_t> stg((size_t)stg_steps * 32, STG_PAD);
  for (int p = 0; p < T1; ++p)
    stg[(size_t)p] = ((uint32_t)s_index[(size_t)p] << 16) | ((uint32_t)stg_q[(size_t)p] << 8) | (uint32_t)stg_e[(size_t)p];
  std::vector<uint32_t> p1((size_t)s1.S * 32), p2((size_t)s2.S * 32);
  std::vector<double> c2((size_t)s2.S * 32, 0.0);
  // a lane whose program has ended keeps stepping: it reads some DATA word of a bank that none of the phase's active
  // lanes uses (no conflict, the value is added to an accumulator that is never flushed again)
  auto idle_word = [&](const Schedule &sc, size_t k, const std::vector<Term> &terms, const std::vector<int> &index_of) {
    const size_t base = k & ~(size_t)15;
    bool used[16] = {false};
    for (size_t j = base; j < base + 16; ++j)
      if (sc.term[j] >= 0) used[index_of[(size_t)terms[(size_t)sc.term[j]].src] & 15] = true;
    for (int idx : index_of)
      if (!used[idx & 15]) return idx;
    return index_of[0];
  };
  for (size_t k = 0; k < p1.size(); ++k) {
    const int t = s1.term[k], f = s1.flush[k];
    const int src = t < 0 ? idle_word(s1, k, t1, s_index) : s_index[(size_t)t1[(size_t)t].src];
    p1[k] = (uint32_t)(8 * src) | ((f < 0 ? NO_FLUSH : (uint32_t)(8 * out1_index(f))) << 16);
  }
  for (size_t k = 0; k < p2.size(); ++k) {
    const int t = s2.term[k], f = s2.flush[k];
    const int src = t < 0 ? idle_word(s2, k, t2, o1_index) : o1_index[(size_t)t2[(size_t)t].src];
    p2[k] = (uint32_t)(8 * src) | ((f < 0 ? NO_FLUSH : (uint32_t)(8 * out2_index(f))) << 16);
    if (t >= 0) c2[k] = t2[(size_t)t].coef;
  }
  // pieces of split destinations are added in ROUNDS: round j adds the (j+1)-th piece of every split destination into
  // it, so the pairs of one round touch distinct destinations (one lane each, no race) and the order is fixed.
  // Section layout: ptr[nr + 1] (in pairs, padded to an even count), then the pairs (destination index, extra index).
  auto rounds_of = [&](const Schedule &sc, int ext, const std::vector<int> &dest_index, int *n_rounds) {
    int nr = 0;
    for (size_t g = 0; g < sc.gd.size(); ++g) nr = std::max(nr, sc.gptr[g + 1] - sc.gptr[g]);
    std::vector<uint16_t> ptr, pairs;
    ptr.push_back(0);
    for (int j = 0; j < nr; ++j) {
      for (size_t g = 0; g < sc.gd.size(); ++g)
        if (sc.gptr[g + 1] - sc.gptr[g] > j) {
          pairs.push_back((uint16_t)dest_index[(size_t)sc.gd[g]]);
          pairs.push_back((uint16_t)(ext + sc.gptr[g] + j));
        }
      ptr.push_back((uint16_t)(pairs.size() / 2));
    }
    *n_rounds = nr;
    while (ptr.size() < (size_t)((nr + 2) & ~1)) ptr.push_back(ptr.back());  // the pairs start 4-byte aligned
    ptr.insert(ptr.end(), pairs.begin(), pairs.end());
    return ptr;
  };
  std::vector<int> o2_index((size_t)r.n2);
  for (int o = 0; o < r.n2; ++o) o2_index[(size_t)o] = 1 + o;
  int nr1 = 0, nr2 = 0;
  std::vector<uint16_t> g1 = rounds_of(s1, ext1, o1_index, &nr1), g2 = rounds_of(s2, ext2, o2_index, &nr2);
  // stage-2 destinations with one or two extra pieces: added while the row is written (no combine pass)
  std::vector<uint32_t> xw((size_t)((r.n2 + 31) / 32) * 32, 0u);
  int wc2 = 1;
  for (size_t g = 0; g < s2.gd.size(); ++g) {
    const int cnt = s2.gptr[g + 1] - s2.gptr[g];
    if (cnt > 2) {
      wc2 = 0;
      break;
    }
    const uint32_t x1 = (uint32_t)(ext2 + s2.gptr[g]), x2 = cnt > 1 ? x1 + 1 : 0u;
    xw[(size_t)s2.gd[g]] = x1 | (x2 << 16);
  }
  Header h;
  memset(&h, 0, sizeof(h));
  h.wc2 = wc2;
  h.n0 = r.n0;
  h.T1 = T1;
  h.stg_steps = stg_steps;
  h.n1 = r.n1;
  h.n2 = r.n2;
  h.S1 = s1.S;
  h.S2 = s2.S;
  h.ng1 = nr1;  // rounds of the combine pass
  h.nx1 = s1.n_extra;
  h.ng2 = nr2;
  h.nx2 = s2.n_extra;
  h.ext1 = ext1;
  h.ext2 = ext2;
  size_t off = align16(sizeof(Header));
  auto place = [&](size_t bytes) {
    size_t at = off;
    off = align16(off + bytes);
    return (int)at;
  };
  h.off_stg = place(stg.size() * 4);
  h.off_w = place((size_t)r.n0 * 8);
  h.off_p1 = place(p1.size() * 4);
  h.off_g1 = place(g1.size() * 2);
  h.off_c2 = place(c2.size() * 8);
  h.off_p2 = place(p2.size() * 4);
  h.off_g2 = place(g2.size() * 2);
  h.off_xw = place(xw.size() * 4);
  h.blob_bytes = (int)off;
  out.blob.assign(off, 0);
  unsigned char *b = out.blob.data();
  memcpy(b, &h, sizeof(h));
  memcpy(b + h.off_stg, stg.data(), stg.size() * 4);
  memcpy(b + h.off_w, r.w.data(), (size_t)r.n0 * 8);
  memcpy(b + h.off_p1, p1.data(), p1.size() * 4);
  if (!g1.empty()) memcpy(b + h.off_g1, g1.data(), g1.size() * 2);
  memcpy(b + h.off_c2, c2.data(), c2.size() * 8);
  memcpy(b + h.off_p2, p2.data(), p2.size() * 4);
  if (!g2.empty()) memcpy(b + h.off_g2, g2.data(), g2.size() * 2);
  memcpy(b + h.off_xw, xw.data(), xw.size() * 4);
  out.n0 = r.n0;
  out.s_cap = s_words;
  out.o1_cap = ext1 + s1.n_extra;
  out.o2_cap = ext2 + s2.n_extra;
  out.use1 = s1.lane_use;
  out.use2 = s2.lane_use;
  out.conf1 = stage_conflicts(s1, t1, s_index);
  out.conf2 = stage_conflicts(s2, t2, o1_index);
  return true;
}

// CPU interpreter of one row: `a_rows[q]` points at the values of operand row q of A_f.  Mirrors
// k_ptap_numeric_tpl (same order of every floating-point operation).
inline void interpret(const unsigned char *blob, const double *const *a_rows, double *c_out) {
  Header h;
  memcpy(&h, blob, sizeof(h));
  const uint32_t *stg = (const uint32_t *)(blob + h.off_stg);
  const double *w = (const double *)(blob + h.off_w);
  const uint32_t *p1 = (const uint32_t *)(blob + h.off_p1);
  const uint16_t *gp1 = (const uint16_t *)(blob + h.off_g1), *pr1 = gp1 + ((h.ng1 + 2) & ~1);
  const double *c2 = (const double *)(blob + h.off_c2);
  const uint32_t *p2 = (const uint32_t *)(blob + h.off_p2);
  const uint16_t *gp2 = (const uint16_t *)(blob + h.off_g2), *pr2 = gp2 + ((h.ng2 + 2) & ~1);
  std::vector<double> S(8192, 0.0), O1((size_t)h.ext1 + h.nx1 + 1, 0.0), O2((size_t)h.ext2 + h.nx2 + 1, 0.0);
  for (int p = 0; p < h.stg_steps * 32; ++p) {
    const uint32_t m = stg[p];
    if (m == STG_PAD) continue;
    const uint32_t q = (m >> 8) & 255u, e = m & 255u;
    S[m >> 16] = w[q] * a_rows[q][e];
  }
  for (int lane = 0; lane < 32; ++lane) {
    double acc = 0.0;
    for (int s = 0; s < h.S1; ++s) {
      const uint32_t u = p1[(size_t)s * 32 + lane];
      acc += S[(u & 0xFFFFu) >> 3];
      const uint32_t f = u >> 16;
      if (f != NO_FLUSH) {
        O1[f >> 3] = acc;
        acc = 0.0;
      }
    }
  }
  for (int rd = 0; rd < h.ng1; ++rd)
    for (int k = gp1[rd]; k < gp1[rd + 1]; ++k) O1[pr1[2 * k]] += O1[pr1[2 * k + 1]];
  for (int lane = 0; lane < 32; ++lane) {
    double acc = 0.0;
    for (int s = 0; s < h.S2; ++s) {
      const uint32_t u = p2[(size_t)s * 32 + lane];
      acc = __builtin_fma(c2[(size_t)s * 32 + lane], O1[(u & 0xFFFFu) >> 3], acc);
      const uint32_t f = u >> 16;
      if (f != NO_FLUSH) {
        O2[f >> 3] = acc;
        acc = 0.0;
      }
    }
  }
  if (h.wc2) {  // the kernel adds the extras while it writes the row: (piece 0 + piece 1) + piece 2, O2[0] = 0
    const uint32_t *xw = (const uint32_t *)(blob + h.off_xw);
    O2[0] = 0.0;
    for (int o = 0; o < h.n2; ++o) c_out[o] = (O2[(size_t)1 + o] + O2[xw[o] & 0xFFFFu]) + O2[xw[o] >> 16];
    return;
  }
  for (int rd = 0; rd < h.ng2; ++rd)
    for (int k = gp2[rd]; k < gp2[rd + 1]; ++k) O2[pr2[2 * k]] += O2[pr2[2 * k + 1]];
  for (int o = 0; o < h.n2; ++o) c_out[o] = O2[(size_t)1 + o];
}

}  // namespace tpl
}  // namespace iife
