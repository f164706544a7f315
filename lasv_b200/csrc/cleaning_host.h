// cleaning_host.h -- the sequential decisions of the reference's graph cleaning
// (`--remove_low_coverage_supernodes T`: tip clipping, then low-coverage supernode
// removal; graph_operations.c:1069-1090), taken on the PATH RECORDS the supernode
// kernels enumerate instead of on the table.
//
// Both passes of the reference walk the table slot by slot and mutate the graph as
// they go, so their result depends on the slot order.  But every walk they make runs
// along whole paths between non-interior nodes, and every mutation either removes
// whole paths or clears one edge bit of a non-interior node.  The graph they see is
// therefore the COMPRESSED graph -- one vertex per non-interior node, one edge per
// path record (length, coverage summary) -- a few per cent of the table, and replaying
// the reference's traversal on it, in slot order, is cheap and exact:
//
//   clip tips      a node with no edges on one side walks out of its other side while
//                  every node has one way on; it is a tip iff within k+1 edges it meets a
//                  node with another way in (dB_graph_population.c:365-441).  Clipping
//                  clears that node's edge bit -- which may let a LATER walk run on
//                  through it.
//   low coverage   a node with 0 < coverage <= T and status none gets its supernode
//                  computed and marked visited; if it has 2..2k+2 edges and no interior
//                  node above T, its interior is pruned and its two end nodes lose the
//                  edge into it -- which may merge the supernodes around them
//                  (:3371-3462).  A path's interior nodes share their fate, so one event
//                  per path (its lowest low-coverage interior slot) and one per
//                  low-coverage non-interior node are all the triggers there are.
//
// The result is a list of actions (prune the interior of a path, prune a node, clear
// edge bits of a node) that the device applies.  Not reproduced: the reference's walk
// limit inside the supernode computation (--max_var_len, 10 000 edges).
//
#pragma once
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "cleaning.cuh"

namespace lasv {

struct ClNode {  // a non-interior node
  uint64_t q;
  uint32_t covg;
  uint8_t edges;
  uint8_t status;  // 1 none, 2 visited, 3 pruned
  int32_t adj[8];  // path leaving through edge bit (side * 4 + nucleotide), -1 if none
};

struct ClPath {
  int32_t a, b;           // end nodes (indices into nodes)
  uint8_t a_bit, b_bit;   // the edge bit of each end node that leads into the path
  uint8_t alive, visited;
  uint32_t len;           // edges
  uint32_t max_covg;      // over the interior nodes
  uint64_t min_low_q;     // lowest interior slot with 0 < coverage <= T, SN_NONE if none
  uint64_t s_q;           // how to walk it: start node, orientation | nucleotide << 1
  uint32_t s_on;
};

struct ClCycle {  // a cycle of interior nodes (record walked forward from its lowest slot s_q)
  uint64_t s_q;
  uint32_t s_on, len;
  uint32_t s_covg, max_covg;  // of the start node / of all the others
  uint64_t min_low_q;         // lowest slot of the whole cycle with 0 < coverage <= T
  uint32_t both_ways;         // crosses every node in both orientations: the node it is printed from is interior too
};

struct ClActions {
  std::vector<ClPathPrune> paths;
  std::vector<uint64_t> nodes;
  std::vector<ClEdgeReset> resets;
  uint64_t tips = 0, tip_edges = 0, supernodes_pruned = 0;
  uint64_t twice = 0;  // interior nodes that a pruning walk meets a second time (paths that turn round)
  // nodes that go from none to pruned when the actions are applied
  uint64_t nodes_pruned() const {
    uint64_t n = nodes.size();
    for (const ClPathPrune &p : paths) n += p.len - 1 - (p.keep_q != SN_NONE ? 1u : 0u);
    return n - twice;
  }
};

class ClGraph {
 public:
  std::vector<ClNode> nodes;  // sorted by slot
  std::vector<ClPath> paths;
  std::vector<ClCycle> cycles;

 private:
  std::vector<uint64_t> bits_;   // one bit per slot: a node of the list sits there
  std::vector<uint32_t> rank_;   // nodes in the slots before each 64-slot word

 public:
  // node list: every occupied, un-pruned, non-interior node with at least one edge, in any order
  void add_node(uint64_t q, uint32_t covg, uint32_t edges) {
    ClNode n;
    n.q = q;
    n.covg = covg;
    n.edges = (uint8_t)edges;
    n.status = 1;
    for (int i = 0; i < 8; i++) n.adj[i] = -1;
    nodes.push_back(n);
  }
  // Put the nodes in slot order and build the slot -> index map: one bit per slot plus a running count
  // per 64 slots, so placing a node and finding the end node of a path are O(1) (the node list arrives
  // in atomic-append order; sorting millions of records and a binary search per path end dominated
  // the host side of a pass otherwise).
  void seal_nodes(uint64_t n_slots) {
    bits_.assign((size_t)((n_slots + 63) / 64), 0ull);
    const std::vector<ClNode> &in = nodes;
    parallel_for(in.size(), [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++)
        if (in[i].q < n_slots) __atomic_fetch_or(&bits_[(size_t)(in[i].q >> 6)], 1ull << (in[i].q & 63), __ATOMIC_RELAXED);
    });
    rank_.assign(bits_.size() + 1, 0u);
    for (size_t w = 0; w < bits_.size(); w++) rank_[w + 1] = rank_[w] + (uint32_t)__builtin_popcountll(bits_[w]);
    std::vector<ClNode> placed(rank_.back());
    parallel_for(in.size(), [&](uint64_t lo, uint64_t hi) {  // every node has a place of its own
      for (uint64_t i = lo; i < hi; i++) {
        const int32_t at = index_of(in[i].q);
        if (at >= 0) placed[(size_t)at] = in[i];
      }
    });
    nodes.swap(placed);
  }
  // ranges of [0, n) on up to 8 host threads (one below parallel_min() items)
  template <class F>
  static void parallel_for(uint64_t n, F &&work) {
    const unsigned threads = n >= parallel_min() ? std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2)) : 1u;
    if (threads <= 1) {
      work(0, n);
      return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++) pool.emplace_back([&work, n, t, threads] { work(n * t / threads, n * (t + 1) / threads); });
    for (auto &th : pool) th.join();
  }
  int32_t index_of(uint64_t q) const {
    const size_t w = (size_t)(q >> 6);
    if (w >= bits_.size()) return -1;
    const uint64_t bit = 1ull << (q & 63);
    if (!(bits_[w] & bit)) return -1;
    return (int32_t)(rank_[w] + (uint32_t)__builtin_popcountll(bits_[w] & (bit - 1)));
  }
  // path record (not a cycle) with its coverage summary; false if an end node is unknown
  bool add_path(const SnPath &p, uint32_t max_covg, uint64_t min_low_q) {
    ClPath c;
    c.a = index_of(p.s_q);
    c.b = index_of(p.t_q);
    if (c.a < 0 || c.b < 0) return false;
    c.a_bit = (uint8_t)(((p.bits & SNB_S_O) ? 4u : 0u) | ((p.bits >> SNB_S_N) & 3u));
    c.b_bit = (uint8_t)(((p.bits & SNB_T_O) ? 4u : 0u) | ((p.bits >> SNB_T_N) & 3u));
    c.alive = 1;
    c.visited = 0;
    c.len = p.len;
    c.max_covg = max_covg;
    c.min_low_q = min_low_q;
    c.s_q = p.s_q;
    c.s_on = ((p.bits & SNB_S_O) ? 1u : 0u) | (((p.bits >> SNB_S_N) & 3u) << 1);
    const int32_t id = (int32_t)paths.size();
    nodes[c.a].adj[c.a_bit] = id;
    nodes[c.b].adj[c.b_bit] = id;
    paths.push_back(c);
    return true;
  }

  // All path records at once (cycles excluded): the ends of a path are looked up in the slot map and its id
  // is written into two adjacency entries that no other path writes (an edge bit of a node leads into exactly
  // one path), so ranges of records go to threads without any locking.  max_covg / min_low_q may be NULL
  // (branch printing needs neither).  False if an end node is unknown.
  static uint64_t &parallel_min() {  // fewer records than this are not worth threads (tests lower it)
    static uint64_t v = 1u << 18;
    return v;
  }
  bool add_paths(const SnPath *recs, const uint32_t *max_covg, const uint64_t *min_low_q, uint64_t n) {
    const size_t base = paths.size();
    paths.resize(base + (size_t)n);
    std::atomic<bool> ok(true);
    auto work = [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) {
        const SnPath &p = recs[i];
        ClPath c;
        c.a = index_of(p.s_q);
        c.b = index_of(p.t_q);
        if (c.a < 0 || c.b < 0) { ok = false; c.a = c.b = 0; }
        c.a_bit = (uint8_t)(((p.bits & SNB_S_O) ? 4u : 0u) | ((p.bits >> SNB_S_N) & 3u));
        c.b_bit = (uint8_t)(((p.bits & SNB_T_O) ? 4u : 0u) | ((p.bits >> SNB_T_N) & 3u));
        c.alive = 1;
        c.visited = 0;
        c.len = p.len;
        c.max_covg = max_covg ? max_covg[i] : 0u;
        c.min_low_q = min_low_q ? min_low_q[i] : SN_NONE;
        c.s_q = p.s_q;
        c.s_on = ((p.bits & SNB_S_O) ? 1u : 0u) | (((p.bits >> SNB_S_N) & 3u) << 1);
        const int32_t id = (int32_t)(base + i);
        if (ok) {
          nodes[(size_t)c.a].adj[c.a_bit] = id;
          nodes[(size_t)c.b].adj[c.b_bit] = id;
        }
        paths[(size_t)id] = c;
      }
    };
    parallel_for(n, work);
    return ok;
  }

  // ------------------------------------------------------------------ helpers
  static uint32_t side_bits(const ClNode &n, uint32_t side) { return (n.edges >> (4u * side)) & 15u; }
  static bool interior_now(const ClNode &n) { return popc4(side_bits(n, 0)) == 1u && popc4(side_bits(n, 1)) == 1u; }
  void other_end(int32_t pid, int32_t node, uint32_t bit, int32_t &node2, uint32_t &bit2) const {
    const ClPath &p = paths[pid];
    if (p.a == node && p.a_bit == bit) { node2 = p.b; bit2 = p.b_bit; }
    else { node2 = p.a; bit2 = p.a_bit; }
  }
  void drop_edge(int32_t node, uint32_t bit, ClActions &act) {
    nodes[node].edges &= (uint8_t)~(1u << bit);
    nodes[node].adj[bit] = -1;
    act.resets.push_back({nodes[node].q, 1u << bit, 0u});
  }
  void prune_node(int32_t node, ClActions &act) {
    nodes[node].status = 3;
    nodes[node].edges = 0;
    act.nodes.push_back(nodes[node].q);
  }
  void prune_path(int32_t pid, uint64_t keep_q, ClActions &act) {
    ClPath &p = paths[pid];
    if (!p.alive) return;
    p.alive = 0;
    if (p.len <= 1) return;
    act.paths.push_back({p.s_q, p.s_on, p.len, keep_q});
    if (p.a == p.b && p.a_bit == p.b_bit) act.twice += (p.len - 1) / 2;  // its own reverse complement: out and back
  }

  // ------------------------------------------------------------------ pass 1
  // db_graph_clip_tips_in_union_of_all_colours (dB_graph_population.c:2563-2608, 445-470, 365-441)
  void clip_tips(int k, ClActions &act) {
    const uint64_t limit = 1 + (uint64_t)k;
    std::vector<int32_t> wn, wp;
    for (int32_t i = 0; i < (int32_t)nodes.size(); i++) {
      if (nodes[i].status != 1) continue;
      for (uint32_t o = 0; o < 2; o++) {  // forward first, reverse only if nothing was clipped
        if (side_bits(nodes[i], o ^ 1u) != 0) continue;  // not a blunt end against the walk
        wn.clear();
        wp.clear();
        int32_t cur = i, y = -1;
        uint32_t side = o, ybit = 0;
        uint64_t length = 0;
        bool join = false;
        for (;;) {
          const uint32_t e = side_bits(nodes[cur], side);
          if (popc4(e) != 1u) break;
          const uint32_t bit = side * 4u + low_bit4(e);
          const int32_t pid = nodes[cur].adj[bit];
          if (pid < 0) break;  // inconsistent input
          other_end(pid, cur, bit, y, ybit);
          wn.push_back(cur);
          wp.push_back(pid);
          length += paths[pid].len;
          if (length > limit) break;
          if (popc4(side_bits(nodes[y], ybit >> 2)) != 1u) { join = true; break; }
          cur = y;
          side = (ybit >> 2) ^ 1u;
        }
        if (!join) continue;
        for (int32_t n : wn) prune_node(n, act);
        for (int32_t p : wp) prune_path(p, SN_NONE, act);
        drop_edge(y, ybit, act);
        act.tips++;
        act.tip_edges += length;
        break;
      }
    }
  }

  // ------------------------------------------------------------------ pass 2
  // db_graph_remove_errors_considering_covg_and_topology (dB_graph_population.c:3595-3649, 3371-3462)
  void remove_low_coverage_supernodes(int k, uint32_t T, ClActions &act) {
    struct Event {
      uint64_t q;
      uint32_t kind;  // 0 path (its lowest low-coverage interior node), 1 node, 2 cycle
      int32_t id;
    };
    std::vector<Event> ev;
    for (int32_t i = 0; i < (int32_t)paths.size(); i++)
      if (paths[i].min_low_q != SN_NONE) ev.push_back({paths[i].min_low_q, 0u, i});
    for (int32_t i = 0; i < (int32_t)nodes.size(); i++)
      if (nodes[i].covg > 0 && nodes[i].covg <= T) ev.push_back({nodes[i].q, 1u, i});
    for (int32_t i = 0; i < (int32_t)cycles.size(); i++)
      if (cycles[i].min_low_q != SN_NONE) ev.push_back({cycles[i].min_low_q, 2u, i});
    std::sort(ev.begin(), ev.end(), [](const Event &x, const Event &y) { return x.q < y.q; });

    std::vector<int32_t> cp, cn;  // paths and interior-now nodes of the supernode at hand
    for (const Event &e : ev) {
      if (e.kind == 2) {  // a cycle of interior nodes: printed from its first low node, which stays
        const ClCycle &c = cycles[e.id];
        if (c.len <= 1 || c.len > 2u * (uint32_t)k + 2u) continue;
        if (c.max_covg > T || c.s_covg > T) continue;
        if (c.both_ways) {  // no node is spared; the walk from s meets every other node twice and s once more
          act.paths.push_back({c.s_q, c.s_on, c.len, SN_NONE});
          act.twice += (c.len - 1) - c.len / 2;
        } else {
          act.paths.push_back({c.s_q, c.s_on, c.len, c.s_q == e.q ? SN_NONE : e.q});
          if (c.s_q == e.q) act.resets.push_back({c.s_q, 0xFFu, 0u});
          else act.nodes.push_back(c.s_q);
        }
        act.supernodes_pruned++;
        continue;
      }
      cp.clear();
      cn.clear();
      int32_t term[2] = {-1, -1};  // end nodes of the supernode and their edge bits into it
      uint32_t term_bit[2] = {0, 0};
      bool cycle = false;
      int32_t cycle_start_node = -1;  // a cycle printed from a (now interior) node: that node stays
      // A path that is its own reverse complement leaves and re-enters its node through the same
      // edge: the reference's walk turns round there and runs back over the chain it came along
      // (every node once more, in the other orientation) to whatever ends the supernode on the
      // OTHER side.  Such a side has no end node of its own; its paths count twice in the length.
      bool reflect[2] = {false, false};
      uint64_t hairpin_len = 0;
      // grow from (node, bit): take the path behind that edge, go on through interior nodes
      auto grow = [&](int32_t node, uint32_t bit, int which, int32_t stop_path, int32_t stop_node) {
        for (;;) {
          const int32_t pid = nodes[node].adj[bit];
          if (pid < 0) { term[which] = node; term_bit[which] = bit; return; }  // inconsistent input
          if (pid == stop_path) { cycle = true; return; }
          cp.push_back(pid);
          int32_t y;
          uint32_t ybit;
          other_end(pid, node, bit, y, ybit);
          if (y == node && ybit == bit) { reflect[which] = true; hairpin_len += paths[pid].len; return; }
          if (y == stop_node) { cycle = true; return; }
          if (!interior_now(nodes[y])) { term[which] = y; term_bit[which] = ybit; return; }
          cn.push_back(y);
          const uint32_t side = (ybit >> 2) ^ 1u;
          node = y;
          bit = side * 4u + low_bit4(side_bits(nodes[y], side));
        }
      };
      // beyond end node `node` of the trigger path (its bit `bit` leads INTO that path)
      auto beyond = [&](int32_t node, uint32_t bit, int which, int32_t stop_path) {
        if (!interior_now(nodes[node])) { term[which] = node; term_bit[which] = bit; return; }
        cn.push_back(node);
        const uint32_t side = (bit >> 2) ^ 1u;
        grow(node, side * 4u + low_bit4(side_bits(nodes[node], side)), which, stop_path, -1);
      };
      uint64_t keep_q = SN_NONE;
      if (e.kind == 0) {
        const ClPath &p = paths[e.id];
        if (!p.alive || p.visited) continue;
        cp.push_back(e.id);
        beyond(p.a, p.a_bit, 0, e.id);
        if (p.a == p.b && p.a_bit == p.b_bit) {  // the path itself turns round: it has one end only
          reflect[1] = true;
          hairpin_len += p.len;
        } else if (!cycle) {
          beyond(p.b, p.b_bit, 1, e.id);
        }
        if (cycle) keep_q = e.q;  // printed from this interior node: everything else is "interior"
      } else {
        ClNode &j = nodes[e.id];
        if (j.status != 1) continue;
        if (interior_now(j)) {
          cycle_start_node = e.id;
          grow(e.id, 0u * 4u + low_bit4(side_bits(j, 0)), 0, -1, e.id);
          if (!cycle) grow(e.id, 1u * 4u + low_bit4(side_bits(j, 1)), 1, -1, e.id);
          if (!cycle) cn.push_back(e.id);
        } else {
          // the reference walks against the node's reverse side first (dB_graph_population.c:1481-1512)
          const uint32_t side = popc4(side_bits(j, 1)) == 1u ? 1u : popc4(side_bits(j, 0)) == 1u ? 0u : 2u;
          if (side == 2u) { j.status = 2; continue; }
          term[0] = e.id;
          term_bit[0] = side * 4u + low_bit4(side_bits(j, side));
          grow(e.id, term_bit[0], 1, -1, -1);
        }
      }
      std::sort(cn.begin(), cn.end());
      cn.erase(std::unique(cn.begin(), cn.end()), cn.end());
      // db_node_action_set_status_visited on every node of the supernode
      for (int32_t p : cp) paths[p].visited = 1;
      for (int32_t n : cn) nodes[n].status = 2;
      if (cycle && cycle_start_node >= 0) nodes[cycle_start_node].status = 2;
      for (int w = 0; w < 2; w++)
        if (term[w] >= 0) nodes[term[w]].status = 2;
      uint64_t length = 0;
      bool all_low = true;
      for (int32_t p : cp) {
        length += paths[p].len;
        all_low = all_low && paths[p].max_covg <= T;
      }
      if (reflect[0] || reflect[1]) {
        length = 2 * (length - hairpin_len) + hairpin_len;
        if (reflect[0] && reflect[1]) {
          // turned round at both ends: a cycle through both orientations of every node, the one it
          // is printed from included -- nothing is spared and there are no end nodes
          cycle = true;
          cycle_start_node = -1;
        }
      }
      for (int32_t n : cn) all_low = all_low && nodes[n].covg <= T;
      if (length <= 1 || length > 2ull * (uint64_t)k + 2ull || !all_low) continue;
      for (int32_t n : cn) prune_node(n, act);
      for (int32_t p : cp) prune_path(p, (e.kind == 0 && p == e.id) ? keep_q : SN_NONE, act);
      if (cycle && cycle_start_node >= 0) {  // the start node stays, without edges
        nodes[cycle_start_node].edges = 0;
        for (int b = 0; b < 8; b++) nodes[cycle_start_node].adj[b] = -1;
        act.resets.push_back({nodes[cycle_start_node].q, 0xFFu, 0u});
      }
      for (int w = 0; w < 2; w++)
        if (term[w] >= 0) drop_edge(term[w], term_bit[w], act);
      act.supernodes_pruned++;
    }
  }
};

}  // namespace lasv
