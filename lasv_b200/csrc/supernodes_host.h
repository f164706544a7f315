// supernodes_host.h -- the sequential part of supernode printing: which of the
// paths the kernels enumerated does the reference print, in which order and from
// which end, for a traversal in slot order (see supernodes.cuh for the argument).
// Plain C++ over the path records; no table access, linear in the number of paths
// after one sort.  Used by capi.cu (product) and by tests/host_emul.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "supernodes.cuh"

namespace lasv {

struct SnReplayStats {
  uint64_t printed = 0;        // supernodes written
  uint64_t by_interior = 0;    // ... triggered by an interior node
  uint64_t by_end_node = 0;    // ... triggered by one of their end nodes
  uint64_t cycles = 0;         // ... that are cycles
  uint64_t never_printed = 0;  // paths without interior nodes whose end nodes were both visited first
  uint64_t longest = 0;        // edges
  uint64_t total_bases = 0;    // sum of k + len over the printed supernodes
};

// The replay proper: feed it the triggers of all paths in slot order (in as many pieces
// as convenient -- the library streams them from the device); `prints` collects the
// printed supernodes in the reference's print order (offsets not filled in).
class SnReplayer {
 public:
  SnReplayer(size_t n_paths, uint64_t n_slots, int k)
      : visited_((n_slots + 63) / 64, 0), printed_(n_paths, 0), n_paths_(n_paths), k_(k) {
    prints.reserve(n_paths);
  }
  void feed(const SnEvent *ev, size_t n) {
    constexpr size_t AHEAD = 16;  // the marks are random accesses into tens of MB: fetch them early
    for (size_t i = 0; i < n; i++) {
      if (i + AHEAD < n) {
        const SnEvent &f = ev[i + AHEAD];
        __builtin_prefetch(&printed_[f.path]);
        __builtin_prefetch(&visited_[f.s_q >> 6]);
        __builtin_prefetch(&visited_[f.t_q >> 6]);
      }
      const SnEvent &e = ev[i];
      if (printed_[e.path]) continue;
      bool from_t = false;  // print the twin (start at the t end)
      if (e.bits & SNB_CYCLE) {
        stats_.cycles++;
      } else if (e.role == 0) {
        // the interior node min_q walks against its own reverse orientation to the end of the
        // supernode and the print starts there: crossed in reverse by the s->t walk => that end is t
        from_t = (e.bits & SNB_MIN_O) != 0;
        stats_.by_interior++;
      } else {
        if (is_visited(e.role == 1 ? e.s_q : e.t_q)) continue;  // an earlier supernode ended here: status != none
        const uint32_t side = (e.bits >> (e.role == 1 ? SNB_S_SIDE : SNB_T_SIDE)) & 3u;
        // reverse-side trigger: walk to the far end, print from there; forward-side: print from x
        from_t = (e.role == 1) == (side == 1u);
        stats_.by_end_node++;
      }
      printed_[e.path] = 1;
      if (!(e.bits & SNB_CYCLE)) {
        visit(e.s_q);
        visit(e.t_q);
      }
      SnPrint pr;
      pr.offset = 0;
      pr.len = e.len;
      if (from_t) {
        pr.q = e.t_q;
        pr.on = ((e.bits & SNB_T_O) ? 1u : 0u) | (((e.bits >> SNB_T_N) & 3u) << 1);
      } else {
        pr.q = e.s_q;
        pr.on = ((e.bits & SNB_S_O) ? 1u : 0u) | (((e.bits >> SNB_S_N) & 3u) << 1);
      }
      prints.push_back(pr);
      if (want_end_keys) {
        // where the printed walk ends: node, orientation it is reached in (the opposite of the twin start there),
        // nucleotide that leads back -- what the per-node writer of ranking.cuh looks its elements up by
        uint64_t key = SN_NONE;
        if (!(e.bits & SNB_CYCLE)) {
          const uint64_t end_q = from_t ? e.s_q : e.t_q;
          const uint32_t twin_o = from_t ? ((e.bits & SNB_S_O) ? 1u : 0u) : ((e.bits & SNB_T_O) ? 1u : 0u);
          const uint32_t back = (e.bits >> (from_t ? SNB_S_N : SNB_T_N)) & 3u;
          key = (end_q << 3) | ((uint64_t)(twin_o ^ 1u) << 2) | back;
        }
        end_keys.push_back(key);
      }
      stats_.printed++;
      stats_.longest = std::max<uint64_t>(stats_.longest, e.len);
      stats_.total_bases += (uint64_t)k_ + e.len;
    }
  }
  SnReplayStats finish() {
    stats_.never_printed = n_paths_ - stats_.printed;
    return stats_;
  }
  std::vector<SnPrint> prints;
  bool want_end_keys = false;      // set before feeding: also collect end_keys[i] for prints[i]
  std::vector<uint64_t> end_keys;

 private:
  bool is_visited(uint64_t q) const { return (visited_[q >> 6] >> (q & 63)) & 1ull; }
  void visit(uint64_t q) { visited_[q >> 6] |= 1ull << (q & 63); }
  // `visited` marks of the nodes that end paths (interior nodes are never asked): one bit per slot
  std::vector<uint64_t> visited_;
  std::vector<uint8_t> printed_;
  size_t n_paths_;
  int k_;
  SnReplayStats stats_;
};

// The same from the path records (path scan, then cycles): builds and sorts the events on
// the host.  tests/host_emul uses this; the library sorts on the device (supernode_kernels.cu).
inline SnReplayStats sn_replay(const std::vector<SnPath> &paths, int k, uint64_t n_slots, std::vector<SnPrint> &out,
                               std::vector<uint64_t> *end_keys = nullptr) {
  struct Keyed {
    uint64_t pos;
    SnEvent e;
  };
  std::vector<Keyed> ev;
  ev.reserve(paths.size() * 2);
  for (size_t i = 0; i < paths.size(); i++) {
    const SnPath &p = paths[i];
    uint64_t pos[3];
    uint32_t role[3];
    const int n = sn_path_events(p, pos, role);
    for (int j = 0; j < n; j++) ev.push_back({pos[j], {p.s_q, p.t_q, p.len, p.bits, (uint32_t)i, role[j]}});
  }
  std::sort(ev.begin(), ev.end(), [](const Keyed &a, const Keyed &b) { return a.pos < b.pos; });
  std::vector<SnEvent> sorted(ev.size());
  for (size_t i = 0; i < ev.size(); i++) sorted[i] = ev[i].e;
  SnReplayer rp(paths.size(), n_slots, k);
  rp.want_end_keys = end_keys != nullptr;
  // in uneven pieces, as the library feeds it
  for (size_t at = 0, step = 1; at < sorted.size(); at += step, step = step * 3 + 1)
    rp.feed(sorted.data() + at, std::min(step, sorted.size() - at));
  out.swap(rp.prints);
  if (end_keys) end_keys->swap(rp.end_keys);
  return rp.finish();
}

// Layout of the reference's FASTA as one file image: record i = header line + sequence
// line.  Sets every SnPrint::offset to where its record starts, returns the image size.
inline uint64_t sn_layout_fasta(std::vector<SnPrint> &prints, int k) {
  uint64_t off = 0;
  for (size_t i = 0; i < prints.size(); i++) {
    prints[i].offset = off;
    const uint64_t bases = (uint64_t)k + prints[i].len;
    off += sn_header_bytes(i, bases, k) + bases + 1;
  }
  return off;
}

}  // namespace lasv
