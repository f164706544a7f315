// branches_host.h -- the sequential part of branch printing (print_branches.c:131-199),
// replayed on the compressed graph of cleaning_host.h: one vertex per non-interior node,
// one edge per path record.
//
// Why O(1) per branch: the walk of print_branches.c:16-127 starts at a neighbour X of a
// branching node J and runs through nodes with one edge each way until a node that is not
// like that, a `visited` node, or 300 further nodes.  So
//   * X is the first interior node of the path behind J's edge, or (path of one edge) the
//     non-interior node at its other end, from which the walk enters the path behind that
//     node's only way on, if it has exactly one;
//   * the nodes it marks on a path are consecutive from the end it entered, and a path is
//     entered at most once from each end (X is marked by the first walk and refused to every
//     later one), so "marked prefix from end a / from end b" is all the state a path needs;
//   * a path that is its own reverse complement turns round in the middle and comes back over
//     its own nodes in the other orientation: position p and position m+1-p are the same node,
//     the walk stops half way at a node it marked itself;
//   * the node that ends the run is added iff the walk is its only way in and it does not have
//     exactly one way on (:112-123) -- whether or not it carries a mark already.
// Cycles of interior nodes are never reached (no branching node leads into them).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "branches.cuh"
#include "cleaning_host.h"

namespace lasv {

struct BrStats {
  uint64_t branches = 0;         // records printed
  uint64_t branching_nodes = 0;  // nodes with more than one edge on a side
  uint64_t nodes = 0;            // sum of the record lengths in nodes
  uint64_t refused = 0;          // neighbours that were marked already
  uint64_t cut = 0;              // walks stopped by the 300-node limit
  uint64_t inconsistent = 0;     // edges without a path record (corrupt input)
  uint64_t image_bytes = 0;
};

class BrReplay {
 public:
  BrReplay(ClGraph &g, int k) : g_(g), k_(k), ma_(g.paths.size(), 0u), mb_(g.paths.size(), 0u) {}

  // every branching node in slot order, sides forward then reverse, nucleotides A C G T
  BrStats run(std::vector<BrPrint> &prints) {
    for (int32_t i = 0; i < (int32_t)g_.nodes.size(); i++) {
      const ClNode &n = g_.nodes[i];
      if (n.covg == 0) continue;  // print_branches.c:146
      const uint32_t d0 = popc4(ClGraph::side_bits(n, 0)), d1 = popc4(ClGraph::side_bits(n, 1));
      if (d0 <= 1 && d1 <= 1) continue;
      st_.branching_nodes++;
      for (uint32_t side = 0; side < 2; side++) {
        const uint32_t e = ClGraph::side_bits(g_.nodes[i], side);
        if (popc4(e) <= 1) continue;
        for (uint32_t nuc = 0; nuc < 4; nuc++)
          if ((e >> nuc) & 1u) branch(i, side * 4u + nuc, prints);
      }
    }
    st_.image_bytes = offset_;
    return st_;
  }

 private:
  ClGraph &g_;
  int k_;
  std::vector<uint32_t> ma_, mb_;  // interior nodes marked from end a / end b of every path
  BrStats st_;
  uint64_t offset_ = 0;

  static bool hairpin(const ClPath &p) { return p.a == p.b && p.a_bit == p.b_bit; }

  // unmarked interior positions of path pid in a row, from position p0 (1 = next to the end entered)
  uint32_t avail(int32_t pid, bool from_a, uint32_t p0) const {
    const ClPath &p = g_.paths[pid];
    const uint32_t m = p.len - 1;
    if (p0 > m) return 0;
    if (hairpin(p)) {
      const uint32_t h = m / 2;
      if (p0 <= ma_[pid] || p0 > h) return 0;
      return h - (p0 - 1);
    }
    const uint32_t near = from_a ? ma_[pid] : mb_[pid], far = from_a ? mb_[pid] : ma_[pid];
    if (p0 <= near) return 0;
    const uint32_t last = m - far;  // last unmarked position
    return last >= p0 ? last - p0 + 1 : 0;
  }
  void mark(int32_t pid, bool from_a, uint32_t upto) {
    uint32_t &near = (from_a || hairpin(g_.paths[pid])) ? ma_[pid] : mb_[pid];
    near = std::max(near, upto);
  }

  void emit(int32_t j, uint32_t bit, uint32_t nodes, std::vector<BrPrint> &prints) {
    BrPrint pr;
    pr.q = g_.nodes[j].q;
    pr.offset = offset_;
    pr.nodes = nodes;
    pr.on = (bit >> 2) | ((bit & 3u) << 1);
    prints.push_back(pr);
    offset_ += br_record_bytes(k_, nodes);
    st_.branches++;
    st_.nodes += nodes;
  }

  void branch(int32_t j, uint32_t bit, std::vector<BrPrint> &prints) {
    const int32_t pid = g_.nodes[j].adj[bit];
    if (pid < 0) { st_.inconsistent++; return; }
    const bool from_a = g_.paths[pid].a == j && g_.paths[pid].a_bit == bit;
    const uint32_t m = g_.paths[pid].len - 1;
    uint32_t length = 1, p0;
    int32_t cur_pid, cur_node;
    uint32_t cur_bit;
    bool cur_from_a;
    if (m >= 1) {  // X = the first interior node of this path
      if (avail(pid, from_a, 1) == 0) { st_.refused++; return; }
      mark(pid, from_a, 1);
      cur_pid = pid; cur_node = j; cur_bit = bit; cur_from_a = from_a; p0 = 2;
    } else {       // X = the node at the other end
      int32_t y;
      uint32_t ybit;
      g_.other_end(pid, j, bit, y, ybit);
      if (g_.nodes[y].status == 2) { st_.refused++; return; }
      g_.nodes[y].status = 2;
      const uint32_t o = (ybit >> 2) ^ 1u, out = ClGraph::side_bits(g_.nodes[y], o);
      if (popc4(out) != 1u) { emit(j, bit, length, prints); return; }  // :50-59
      cur_node = y;
      cur_bit = o * 4u + low_bit4(out);
      cur_pid = g_.nodes[y].adj[cur_bit];
      if (cur_pid < 0) { st_.inconsistent++; return; }
      cur_from_a = g_.paths[cur_pid].a == y && g_.paths[cur_pid].a_bit == cur_bit;
      p0 = 1;
    }
    // :82-110, `length <= 300` is tested before a node is added
    const uint32_t free_run = avail(cur_pid, cur_from_a, p0), room = BR_LIMIT + 1 - length;
    const uint32_t c = std::min(free_run, room);
    if (c) mark(cur_pid, cur_from_a, p0 - 1 + c);
    length += c;
    const uint32_t mc = g_.paths[cur_pid].len - 1;
    if (p0 - 1 + c == mc && !(hairpin(g_.paths[cur_pid]) && mc > 0)) {  // the run ended at a non-interior node: :112-123
      int32_t y;
      uint32_t ybit;
      g_.other_end(cur_pid, cur_node, cur_bit, y, ybit);
      const uint32_t in = popc4(ClGraph::side_bits(g_.nodes[y], ybit >> 2));
      const uint32_t out = popc4(ClGraph::side_bits(g_.nodes[y], (ybit >> 2) ^ 1u));
      if (in == 1u && out != 1u) {
        g_.nodes[y].status = 2;
        length++;
      }
    } else if (c == room && free_run > room) {
      st_.cut++;
    }
    emit(j, bit, length, prints);
  }
};

}  // namespace lasv
