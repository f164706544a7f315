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

// events: the triggers of all paths in slot order.  out: the printed supernodes in the
// reference's print order (offsets not filled in).
inline SnReplayStats sn_replay_events(const SnEvent *ev, size_t n_events, size_t n_paths, uint64_t n_slots, int k,
                                      std::vector<SnPrint> &out) {
  // `visited` marks of the nodes that end paths (interior nodes are never asked): one bit per slot
  std::vector<uint64_t> visited((n_slots + 63) / 64, 0);
  std::vector<uint8_t> printed(n_paths, 0);
  auto is_visited = [&](uint64_t q) { return (visited[q >> 6] >> (q & 63)) & 1ull; };
  auto visit = [&](uint64_t q) { visited[q >> 6] |= 1ull << (q & 63); };
  SnReplayStats st;
  out.clear();
  out.reserve(n_paths);
  for (size_t i = 0; i < n_events; i++) {
    const SnEvent &e = ev[i];
    if (printed[e.path]) continue;
    bool from_t = false;  // print the twin (start at the t end)
    if (e.bits & SNB_CYCLE) {
      st.cycles++;
    } else if (e.role == 0) {
      // the interior node min_q walks against its own reverse orientation to the end of the
      // supernode and the print starts there: crossed in reverse by the s->t walk => that end is t
      from_t = (e.bits & SNB_MIN_O) != 0;
      st.by_interior++;
    } else {
      if (is_visited(e.role == 1 ? e.s_q : e.t_q)) continue;  // an earlier supernode ended here: status != none
      const uint32_t side = (e.bits >> (e.role == 1 ? SNB_S_SIDE : SNB_T_SIDE)) & 3u;
      // reverse-side trigger: walk to the far end, print from there; forward-side: print from x
      from_t = (e.role == 1) == (side == 1u);
      st.by_end_node++;
    }
    printed[e.path] = 1;
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
    out.push_back(pr);
    st.printed++;
    st.longest = std::max<uint64_t>(st.longest, e.len);
    st.total_bases += (uint64_t)k + e.len;
  }
  st.never_printed = n_paths - st.printed;
  return st;
}

// The same from the path records (path scan, then cycles): builds and sorts the events on
// the host.  tests/host_emul uses this; the library sorts on the device (supernode_kernels.cu).
inline SnReplayStats sn_replay(const std::vector<SnPath> &paths, int k, uint64_t n_slots, std::vector<SnPrint> &out) {
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
  return sn_replay_events(sorted.data(), sorted.size(), paths.size(), n_slots, k, out);
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
