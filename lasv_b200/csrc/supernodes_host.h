// supernodes_host.h -- the sequential part of supernode printing: which of the
// paths the kernels enumerated does the reference print, in which order and from
// which end, for a traversal in slot order (see supernodes.cuh for the argument).
// Plain C++ over the path records; no table access, linear in the number of paths
// after one sort.  Used by capi.cu (product) and by tests/host_emul.
#pragma once
#include <stdint.h>
#include <stdio.h>

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

// paths: records of the path scan followed by the cycle records.  out: the printed
// supernodes in the reference's print order (offsets not filled in).
inline SnReplayStats sn_replay(const std::vector<SnPath> &paths, int k, std::vector<SnPrint> &out) {
  struct Event {
    uint64_t pos;   // slot at which the traversal fires it
    uint32_t path;  // index into paths
    uint32_t role;  // 0 interior node / cycle, 1 start node, 2 end node
  };
  std::vector<Event> ev;
  ev.reserve(paths.size() * 2);
  std::vector<uint64_t> ends;  // the non-interior nodes that end some path: their `visited` marks matter
  ends.reserve(paths.size() * 2);
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
    ends.push_back(p.s_q);
    ends.push_back(p.t_q);
  }
  std::sort(ends.begin(), ends.end());
  ends.erase(std::unique(ends.begin(), ends.end()), ends.end());
  std::sort(ev.begin(), ev.end(), [](const Event &a, const Event &b) {
    return a.pos != b.pos ? a.pos < b.pos : a.role < b.role;
  });
  std::vector<uint8_t> visited(ends.size(), 0), printed(paths.size(), 0);
  auto end_index = [&](uint64_t q) { return (size_t)(std::lower_bound(ends.begin(), ends.end(), q) - ends.begin()); };

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
      if (visited[end_index(x)]) continue;  // an earlier supernode ended here: status != none
      const uint32_t side = (p.bits >> (e.role == 1 ? SNB_S_SIDE : SNB_T_SIDE)) & 3u;
      // reverse-side trigger: walk to the far end, print from there; forward-side: print from x
      from_t = (e.role == 1) == (side == 1u);
      st.by_end_node++;
    }
    printed[e.path] = 1;
    if (!(p.bits & SNB_CYCLE)) {
      visited[end_index(p.s_q)] = 1;
      visited[end_index(p.t_q)] = 1;
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

// Byte offsets of the sequences in one packed buffer; returns its size.
inline uint64_t sn_assign_offsets(std::vector<SnPrint> &prints, int k) {
  uint64_t off = 0;
  for (SnPrint &p : prints) {
    p.offset = off;
    off += (uint64_t)k + p.len;
  }
  return off;
}

// The reference's FASTA (print_ultra_minimal_fasta_from_path with the first k-mer,
// dB_graph_population.c:6167-6169; names node_0, node_1, ... in print order, :2761).
inline bool sn_write_fasta(FILE *fp, const std::vector<SnPrint> &prints, const char *seqs, int k) {
  for (size_t i = 0; i < prints.size(); i++) {
    const SnPrint &p = prints[i];
    const size_t n = (size_t)k + p.len;
    if (fprintf(fp, ">node_%zu length:%zu INFO:KMER=%d\n", i, n, k) < 0) return false;
    if (fwrite(seqs + p.offset, 1, n, fp) != n) return false;
    if (fputc('\n', fp) == EOF) return false;
  }
  return true;
}

}  // namespace lasv
