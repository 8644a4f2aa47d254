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
//   staging     S[1+p] = w[q_p] * A_f.val[beg[q_p] + e_p]       p over the T1 stage-1 terms (coalesced by operand row)
//   stage 1     O1[f] = sum of S[src] over a lane's run of terms  (the intermediate row (R A_f)[i,:], never in HBM)
//   stage 2     O2[f] = sum of coef * O1[src]                     (coef = the M value of the term)
//   write       A_b.val[row i] = O2[1..n2]
//
// Stages 1 and 2 are GATHERS: the terms of one destination are consecutive in one lane's program, accumulate in a
// register and are stored once ("flush"); a destination with more terms than the step count S is split into
// pieces that flush into extra slots and are added in a fixed order afterwards.  No shared-memory read-modify-write,
// no privatised accumulator copies, no hashing, no column indices, no atomics; the sum order is fixed by the
// program, so results are bit-reproducible.  Index 0 of S / O1 is a zero slot read by padding steps.
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
constexpr int MAX_T1 = 1023;   // stage-1 product terms  (S index fits 13 bits after the byte scaling)
constexpr int MAX_N1 = 256;    // intermediate row entries
constexpr int MAX_N2 = 256;    // output row entries
constexpr int MAX_T2 = 4095;   // stage-2 product terms
constexpr int MAX_OUT = 1023;  // destinations + extra slots of a stage (byte offset must fit 16 bits)
constexpr uint32_t NO_FLUSH = 0xFFFFu;
constexpr uint16_t STG_PAD = 0xFFFFu;

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
  int ng1, nx1, ng2, nx2;
  int off_stg, off_w, off_p1, off_g1, off_c2, off_p2, off_g2;
  int blob_bytes, pad;
};
static_assert(sizeof(Header) == 80, "Header layout");

struct Term {
  int dest;     // destination slot 0..n_dest-1
  int src;      // index into the source buffer (already +1: 0 is the zero slot)
  double coef;  // stage 2 only
};

struct Packed {
  int S = 0, n_extra = 0;
  std::vector<uint32_t> prog;  // [S*32]  low 16 bits: BYTE offset of the source entry; high 16: BYTE offset of the flush slot or NO_FLUSH
  std::vector<double> coef;    // [S*32]  (stage 2)
  std::vector<uint16_t> gd;    // [ng]    index (not bytes) of a split destination in the out buffer
  std::vector<uint16_t> gptr;  // [ng+1]  its extras are out[1 + n_dest + gptr[g] .. gptr[g+1])
  double lane_use = 0.0;       // terms / (32 S)
  double wavefronts = 0.0;     // mean shared-memory phases per half-warp read (1.0 = conflict free)
};

// Bin-pack the destinations' term runs into 32 lane programs of S steps, S as small as possible.
// Terms of a destination keep their given order (= ascending operand order, the order the reference's
// row-wise Gustavson product adds them in).  Returns false if a destination has no term or the out buffer
// would exceed MAX_OUT entries.
inline bool pack_stage(int n_dest, const std::vector<Term> &terms, bool with_coef, Packed &out) {
  std::vector<std::vector<int>> by_dest((size_t)n_dest);
  for (size_t t = 0; t < terms.size(); ++t) {
    if (terms[t].dest < 0 || terms[t].dest >= n_dest) return false;
    by_dest[(size_t)terms[t].dest].push_back((int)t);
  }
  for (int d = 0; d < n_dest; ++d)
    if (by_dest[(size_t)d].empty()) return false;
  const int N = (int)terms.size();
  struct Piece {
    int dest, first, count, index;  // terms by_dest[dest][first .. first+count), index-th piece of dest
  };
  std::vector<Piece> pieces;
  std::vector<int> lane_of;  // per piece
  int S = std::max(1, (N + 31) / 32);
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
        out.gd.push_back((uint16_t)(1 + pieces[k].dest));
        out.gptr.push_back((uint16_t)n_extra);
      }
      k = k2;
    }
  }
  if (1 + n_dest + n_extra > MAX_OUT) return false;
  out.S = S;
  out.n_extra = n_extra;
  out.prog.assign((size_t)S * 32, (NO_FLUSH << 16));
  out.coef.assign(with_coef ? (size_t)S * 32 : 0, 0.0);
  // ---- schedule: which term every lane reads at every step.  The 32 lanes of a step read 32 fp64 words of shared
  // memory; the access is served in two phases of 16 lanes, each conflict-free iff its words lie in 16 distinct
  // 8-byte banks (index mod 16) or coincide.  The order of the terms inside a run and the order of a lane's runs are
  // free (any fixed order is a valid, reproducible sum), so a greedy list scheduler picks, lane by lane (most
  // constrained first), a term whose bank is still unused in this phase of this step.
  std::vector<std::vector<int>> lane_pieces(32);
  for (size_t k = 0; k < pieces.size(); ++k) lane_pieces[(size_t)lane_of[k]].push_back((int)k);
  std::vector<std::vector<int>> left(pieces.size());  // remaining term indices of every piece
  for (size_t k = 0; k < pieces.size(); ++k)
    for (int e = 0; e < pieces[k].count; ++e) left[k].push_back(by_dest[(size_t)pieces[k].dest][(size_t)(pieces[k].first + e)]);
  int cur[32];
  for (int l = 0; l < 32; ++l) cur[l] = -1;
  long long wavefronts = 0;
  for (int s = 0; s < S; ++s) {
    for (int half = 0; half < 2; ++half) {
      int load[16] = {0};  // lanes scheduled per bank
      int order[16], n_opt[16];
      bool idle_any = false;
      for (int k = 0; k < 16; ++k) {
        const int l = half * 16 + k;
        order[k] = l;
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
            const int b = terms[(size_t)rem[z]].src & 15;
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
      // maximum matching (Kuhn's augmenting paths), most constrained lanes first; bank 0 is taken when a lane idles
      int bank_owner[16], lane_bank[16];
      for (int b = 0; b < 16; ++b) bank_owner[b] = -1;
      for (int k = 0; k < 16; ++k) lane_bank[k] = -1;
      if (idle_any) bank_owner[0] = 16;  // sentinel: never re-routed
      {
        int n_banks[16];
        for (int k = 0; k < 16; ++k) {
          n_banks[k] = 0;
          for (int b = 0; b < 16; ++b) n_banks[k] += cand_piece[k][b] >= 0;
          order[k] = k;
        }
        std::stable_sort(order, order + 16, [&](int x, int y) { return n_banks[x] < n_banks[y]; });
        for (int oi = 0; oi < 16; ++oi) {
          const int k0 = order[oi];
          if (n_opt[k0] == 0) continue;
          bool visited[16] = {false};
          // iterative DFS would do; recursion depth is at most 16
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
        const int best_piece = cand_piece[k][b], best_pos = cand_pos[k][b];
        std::vector<int> &rem = left[(size_t)best_piece];
        const Term &tm = terms[(size_t)rem[(size_t)best_pos]];
        rem.erase(rem.begin() + best_pos);
        uint32_t word = (uint32_t)(tm.src * 8) & 0xFFFFu;
        uint32_t flush = NO_FLUSH;
        if (rem.empty()) {
          const Piece &pc = pieces[(size_t)best_piece];
          flush = (uint32_t)(8 * (pc.index == 0 ? 1 + pc.dest : 1 + n_dest + extra_id[(size_t)best_piece]));
          cur[l] = -1;
        } else {
          cur[l] = best_piece;
        }
        const size_t at = (size_t)s * 32 + (size_t)l;
        out.prog[at] = word | (flush << 16);
        if (with_coef) out.coef[at] = tm.coef;
      }
      int mx = 1;
      for (int b = 0; b < 16; ++b) mx = std::max(mx, load[b]);
      wavefronts += mx;
    }
  }
  out.wavefronts = (double)wavefronts / (2.0 * S);  // 1.0 = conflict free
  out.lane_use = (double)N / (32.0 * S);
  return true;
}

struct Program {
  std::vector<unsigned char> blob;  // Header + sections
  int s_cap = 0, o1_cap = 0, o2_cap = 0;  // entries of S / O1 / O2 this template needs
  double use1 = 0.0, use2 = 0.0;
  double conf1 = 0.0, conf2 = 0.0;  // mean conflict degree of the gather reads of the two stages (1.0 = none)
};

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

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
  // staging list and stage-1 terms share the term order (operand rows ascending, entries ascending)
  const int stg_steps = (T1 + 31) / 32;
  std::vector<uint16_t> stg((size_t)stg_steps * 32, STG_PAD);
  std::vector<Term> t1((size_t)T1), t2((size_t)T2);
  {
    int p = 0;
    for (int q = 0; q < r.n0; ++q)
      for (int e = 0; e < r.len1[(size_t)q]; ++e, ++p) {
        stg[(size_t)p] = (uint16_t)((q << 8) | e);
        t1[(size_t)p] = {(int)r.slot1[(size_t)p], 1 + p, 0.0};
      }
    p = 0;
    for (int q = 0; q < r.n1; ++q)
      for (int e = 0; e < r.len2[(size_t)q]; ++e, ++p) t2[(size_t)p] = {(int)r.slot2[(size_t)p], 1 + q, r.mval[(size_t)p]};
  }
  Packed p1, p2;
  if (!pack_stage(r.n1, t1, false, p1)) return false;
  if (!pack_stage(r.n2, t2, true, p2)) return false;
  Header h;
  memset(&h, 0, sizeof(h));
  h.n0 = r.n0;
  h.T1 = T1;
  h.stg_steps = stg_steps;
  h.n1 = r.n1;
  h.n2 = r.n2;
  h.S1 = p1.S;
  h.S2 = p2.S;
  h.ng1 = (int)p1.gd.size();
  h.nx1 = p1.n_extra;
  h.ng2 = (int)p2.gd.size();
  h.nx2 = p2.n_extra;
  size_t off = align16(sizeof(Header));
  auto place = [&](size_t bytes) {
    size_t at = off;
    off = align16(off + bytes);
    return (int)at;
  };
  h.off_stg = place(stg.size() * 2);
  h.off_w = place((size_t)r.n0 * 8);
  h.off_p1 = place(p1.prog.size() * 4);
  h.off_g1 = place((p1.gd.size() + p1.gptr.size()) * 2);
  h.off_c2 = place(p2.coef.size() * 8);
  h.off_p2 = place(p2.prog.size() * 4);
  h.off_g2 = place((p2.gd.size() + p2.gptr.size()) * 2);
  h.blob_bytes = (int)off;
  out.blob.assign(off, 0);
  unsigned char *b = out.blob.data();
  memcpy(b, &h, sizeof(h));
  memcpy(b + h.off_stg, stg.data(), stg.size() * 2);
  memcpy(b + h.off_w, r.w.data(), (size_t)r.n0 * 8);
  memcpy(b + h.off_p1, p1.prog.data(), p1.prog.size() * 4);
  if (!p1.gd.empty()) memcpy(b + h.off_g1, p1.gd.data(), p1.gd.size() * 2);
  memcpy(b + h.off_g1 + p1.gd.size() * 2, p1.gptr.data(), p1.gptr.size() * 2);
  memcpy(b + h.off_c2, p2.coef.data(), p2.coef.size() * 8);
  memcpy(b + h.off_p2, p2.prog.data(), p2.prog.size() * 4);
  if (!p2.gd.empty()) memcpy(b + h.off_g2, p2.gd.data(), p2.gd.size() * 2);
  memcpy(b + h.off_g2 + p2.gd.size() * 2, p2.gptr.data(), p2.gptr.size() * 2);
  out.s_cap = 1 + stg_steps * 32;
  out.o1_cap = 1 + r.n1 + p1.n_extra;
  out.o2_cap = 1 + r.n2 + p2.n_extra;
  out.use1 = p1.lane_use;
  out.use2 = p2.lane_use;
  out.conf1 = p1.wavefronts;
  out.conf2 = p2.wavefronts;
  return true;
}

// CPU interpreter of one row: `a_rows[q]` points at the values of operand row q of A_f.  Mirrors
// k_ptap_numeric_tpl (same order of every floating-point operation).
inline void interpret(const unsigned char *blob, const double *const *a_rows, double *c_out) {
  Header h;
  memcpy(&h, blob, sizeof(h));
  const uint16_t *stg = (const uint16_t *)(blob + h.off_stg);
  const double *w = (const double *)(blob + h.off_w);
  const uint32_t *p1 = (const uint32_t *)(blob + h.off_p1);
  const uint16_t *gd1 = (const uint16_t *)(blob + h.off_g1), *gp1 = gd1 + h.ng1;
  const double *c2 = (const double *)(blob + h.off_c2);
  const uint32_t *p2 = (const uint32_t *)(blob + h.off_p2);
  const uint16_t *gd2 = (const uint16_t *)(blob + h.off_g2), *gp2 = gd2 + h.ng2;
  std::vector<double> S((size_t)1 + (size_t)h.stg_steps * 32, 0.0), O1((size_t)1 + h.n1 + h.nx1, 0.0), O2((size_t)1 + h.n2 + h.nx2, 0.0);
  for (int p = 0; p < h.stg_steps * 32; ++p) {
    const uint16_t m = stg[p];
    if (m == STG_PAD) continue;
    S[(size_t)1 + p] = w[m >> 8] * a_rows[m >> 8][m & 255];
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
  for (int g = 0; g < h.ng1; ++g) {
    double v = O1[gd1[g]];
    for (int x = gp1[g]; x < gp1[g + 1]; ++x) v += O1[(size_t)1 + h.n1 + x];
    O1[gd1[g]] = v;
  }
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
  for (int g = 0; g < h.ng2; ++g) {
    double v = O2[gd2[g]];
    for (int x = gp2[g]; x < gp2[g + 1]; ++x) v += O2[(size_t)1 + h.n2 + x];
    O2[gd2[g]] = v;
  }
  for (int o = 0; o < h.n2; ++o) c_out[o] = O2[(size_t)1 + o];
}

}  // namespace tpl
}  // namespace iife
