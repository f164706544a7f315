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

// paths: records of the path scan followed by the cycle records; n_slots: size of the
// table (slot numbers are < n_slots).  out: the printed supernodes in the reference's
// print order (offsets not filled in).
inline SnReplayStats sn_replay(const std::vector<SnPath> &paths, int k, uint64_t n_slots, std::vector<SnPrint> &out) {
  struct Event {
    uint64_t pos;   // slot at which the traversal fires it
    uint32_t path;  // index into paths
    uint32_t role;  // 0 interior node / cycle, 1 start node, 2 end node
  };
  std::vector<Event> ev;
  ev.reserve(paths.size() * 2);
  for (size_t i = 0; i < paths.size(); i++) {
    const SnPath &p = paths[i];
    if (p.bits & SNB_CYCLE) {
      ev.push_back({p.min_q, (uint32_t)i, 0u});
      continue;
    }
    if (p.min_q != SN_NONE) ev.push_back({p.min_q, (uint32_t)i, 0u});
    if (p.bits & SNB_TRIG_S) ev.push_back({p.s_q, (uint32_t)i, 1u});
    // a path that is its own twin carries both flags for the same (node, side): one event
    const bool self_twin = p.s_q == p.t_q && ((p.bits & SNB_S_O) != 0) == ((p.bits & SNB_T_O) != 0);
    if ((p.bits & SNB_TRIG_T) && !self_twin) ev.push_back({p.t_q, (uint32_t)i, 2u});
  }
  std::sort(ev.begin(), ev.end(), [](const Event &a, const Event &b) {
    return a.pos != b.pos ? a.pos < b.pos : a.role < b.role;
  });
  // `visited` marks of the nodes that end paths (interior nodes are never asked): one bit per slot
  std::vector<uint64_t> visited((n_slots + 63) / 64, 0);
  std::vector<uint8_t> printed(paths.size(), 0);
  auto is_visited = [&](uint64_t q) { return (visited[q >> 6] >> (q & 63)) & 1ull; };
  auto visit = [&](uint64_t q) { visited[q >> 6] |= 1ull << (q & 63); };

  SnReplayStats st;
  out.clear();
  for (const Event &e : ev) {
    const SnPath &p = paths[e.path];
    if (printed[e.path]) continue;
    bool from_t = false;  // print the twin (start at the t end)
    if (p.bits & SNB_CYCLE) {
      st.cycles++;
    } else if (e.role == 0) {
      // the interior node min_q walks against its own reverse orientation to the end of the
      // supernode and the print starts there: crossed in reverse by the s->t walk => that end is t
      from_t = (p.bits & SNB_MIN_O) != 0;
      st.by_interior++;
    } else {
      const uint64_t x = e.role == 1 ? p.s_q : p.t_q;
      if (is_visited(x)) continue;  // an earlier supernode ended here: status != none
      const uint32_t side = (p.bits >> (e.role == 1 ? SNB_S_SIDE : SNB_T_SIDE)) & 3u;
      // reverse-side trigger: walk to the far end, print from there; forward-side: print from x
      from_t = (e.role == 1) == (side == 1u);
      st.by_end_node++;
    }
    printed[e.path] = 1;
    if (!(p.bits & SNB_CYCLE)) {
      visit(p.s_q);
      visit(p.t_q);
    }
    SnPrint pr;
    pr.offset = 0;
    pr.len = p.len;
    if (from_t) {
      pr.q = p.t_q;
      pr.on = ((p.bits & SNB_T_O) ? 1u : 0u) | (((p.bits >> SNB_T_N) & 3u) << 1);
    } else {
      pr.q = p.s_q;
      pr.on = ((p.bits & SNB_S_O) ? 1u : 0u) | (((p.bits >> SNB_S_N) & 3u) << 1);
    }
    out.push_back(pr);
    st.printed++;
    st.longest = std::max<uint64_t>(st.longest, p.len);
    st.total_bases += (uint64_t)k + p.len;
  }
  for (size_t i = 0; i < paths.size(); i++) st.never_printed += !printed[i];
  return st;
}

// Layout of the reference's FASTA as one file image: per supernode a header line
// ">node_<i> length:<bases> INFO:KMER=<k>" (print_ultra_minimal_fasta_from_path with the
// first k-mer, dB_graph_population.c:6167-6169; names in print order, :2761) followed by
// the sequence line.  Sets every SnPrint::offset to where its sequence starts (the device
// writes sequence + newline there), returns the image size; hdr_at[i] = where header i starts.
inline uint64_t sn_layout_fasta(std::vector<SnPrint> &prints, int k, std::vector<uint64_t> &hdr_at) {
  hdr_at.resize(prints.size());
  uint64_t off = 0;
  char tmp[96];
  for (size_t i = 0; i < prints.size(); i++) {
    hdr_at[i] = off;
    off += (uint64_t)snprintf(tmp, sizeof tmp, ">node_%zu length:%zu INFO:KMER=%d\n", i, (size_t)k + prints[i].len, k);
    prints[i].offset = off;
    off += (uint64_t)k + prints[i].len + 1;
  }
  return off;
}

inline void sn_fill_headers(char *image, const std::vector<SnPrint> &prints, int k, const std::vector<uint64_t> &hdr_at) {
  char tmp[96];
  for (size_t i = 0; i < prints.size(); i++) {
    const int n = snprintf(tmp, sizeof tmp, ">node_%zu length:%zu INFO:KMER=%d\n", i, (size_t)k + prints[i].len, k);
    memcpy(image + hdr_at[i], tmp, (size_t)n);
  }
}

}  // namespace lasv
