// supernodes.cuh -- maximal unambiguous contigs ("supernodes") of the graph that
// sits in the device table: the per-node / per-path arithmetic shared by the
// kernels (graph_kernels.cu) and by tests/host_emul.
//
// Replaces, for `--output_supernodes` (reference file:line under /root/reference):
//   db_graph_print_supernodes_defined_by_func_of_colours   src/dB_graph_operations/dB_graph_population.c:2696-2803
//   db_graph_supernode_in_subgraph_...                     dB_graph_population.c:1454-1527
//   db_graph_get_perfect_path_with_first_edge_...          dB_graph_population.c:783-896
//   db_graph_get_next_node_in_subgraph_...                 dB_graph_population.c:156-206
//   db_node_has_precisely_one_edge_...                     src/dB_graph_operations/element.c:653-679
//   print_ultra_minimal_fasta_from_path                    dB_graph_population.c:6125-6178
//
// The reference walks the table slot by slot on one thread: a node whose status
// is still `none` gets its supernode computed (walk against the node's reverse
// orientation to the end, then forward from there), printed, and all its nodes
// marked `visited`.  What that prints is a function of the graph and of the
// traversal ORDER only, and it decomposes:
//
//   * A node is INTERIOR iff it has exactly one edge on each side.  Walking, a
//     path continues through a node iff that node is interior (one way in, one
//     way out), so every edge (S, o) --n--> of a non-interior node S starts one
//     path that runs through interior nodes to the next non-interior node T.
//     Each path has a twin read from the other end (its reverse complement).
//     Components made of interior nodes only are cycles.
//   * An interior node belongs to exactly one path and triggers it when the
//     traversal reaches it.  A non-interior node X triggers at most ONE of the
//     paths it ends: the one through its only reverse-side edge, else the one
//     through its only forward-side edge (the reference tries the reverse walk
//     first) -- and only if no path that ends in X was printed before the
//     traversal reached X.  That last condition is the only order dependence,
//     and it is decided by replaying, in slot order, one event per path (its
//     first interior slot) and one per triggering end node (supernodes_host.h).
//
// So: the kernels enumerate all paths in parallel (one walk per path start, both
// twins, the smaller start state keeps the record), the host replays the events
// (linear in the number of paths, no table access), and a last kernel writes
// the sequences of the printed paths.  Output: the reference's FASTA, byte for
// byte, for the same slot order -- which is the order of the `.ctx` we dump, so
// it is what the reference itself prints after reloading that `.ctx`.
// Not reproduced: the reference cuts a walk after `max_length` edges (10 000 by
// default, cmd_line.c:365) and then prints overlapping pieces of the longer path;
// here a path is always whole and the number of such paths is reported.
#pragma once
#include "table.cuh"

namespace lasv {

constexpr uint64_t SN_NONE = ~0ull;

// ------------------------------------------------------------ k-mer arithmetic
LASV_HD uint64_t rev_base_pairs64(uint64_t x) {  // reverse the order of the 32 two-bit groups
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
  return (x >> 32) | (x << 32);
}

// binary_kmer_reverse_complement (binary_kmer.c:456-523) in closed form: complement,
// reverse all 64N bits base-wise, shift the k bases back down to the right end.
template <int N>
LASV_HD void kmer_revcomp(const uint64_t (&in)[N], int k, uint64_t (&out)[N]) {
  uint64_t full[N];
#pragma unroll
  for (int i = 0; i < N; i++) full[i] = rev_base_pairs64(~in[N - 1 - i]);
  const uint32_t s = 64u * N - 2u * (uint32_t)k;  // 2..62: k is odd and 2k/64 + 1 == N
#pragma unroll
  for (int i = N - 1; i >= 0; i--) out[i] = (full[i] >> s) | (i > 0 ? full[i - 1] << (64u - s) : 0ull);
}

// binary_kmer_left_shift_one_base_and_insert_new_base_at_right_end (binary_kmer.c:134-143)
template <int N>
LASV_HD void kmer_shift_append(uint64_t (&km)[N], uint32_t nuc, int k) {
#pragma unroll
  for (int i = 0; i < N - 1; i++) km[i] = (km[i] << 2) | (km[i + 1] >> 62);
  km[N - 1] = (km[N - 1] << 2) | (uint64_t)nuc;
  km[0] &= (~0ull) >> (64 - 2 * (k & 31));
}

template <int N>
LASV_HD uint32_t kmer_first_base(const uint64_t (&km)[N], int k) {
  return (uint32_t)(km[0] >> (2 * (k & 31) - 2)) & 3u;
}

// base i (0 = first) of a k-mer
template <int N>
LASV_HD uint32_t kmer_base_at(const uint64_t (&km)[N], int k, int i) {
  const int bit = 2 * (k - 1 - i);  // from the right end
  return (uint32_t)(km[N - 1 - (bit >> 6)] >> (bit & 63)) & 3u;
}

// ---------------------------------------------------------------- node access
// Nodes are named by their slot number q = bucket * W + s in table order (the
// order of hash_table_traverse and of the .ctx records), in either arrangement.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD uint64_t sn_load_node(const TableView &t, uint64_t q, uint64_t (&key)[N]) {
  const uint32_t bucket = (uint32_t)(q / t.W);
  uint64_t *slot = slot_ptr<N>(t, bucket, (uint32_t)(q - (uint64_t)bucket * t.W));
  const uint64_t m = Ops::load(slot + N);
  if (meta_status(m) == ST_EMPTY) return 0ull;
#pragma unroll
  for (int w = 0; w < N; w++) key[w] = Ops::load(slot + w);
  return m;
}

// hash_table_find returning the slot number (SN_NONE if absent) and the meta word
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD uint64_t sn_locate(const TableView &t, const uint64_t (&key)[N], uint64_t &meta_out) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT, G = LineLayout<N>::G;
  if (t.finalized) {
    for (uint32_t r = 0; r <= t.max_rehash; ++r) {
      const uint32_t b = lookup3_key<N>(key, r).c & t.bucket_mask;
      uint8_t *bucket = bucket_base(t, b);
      for (uint32_t s = 0; s < t.W; ++s) {
        uint64_t *slot = reinterpret_cast<uint64_t *>(bucket + (uint64_t)s * SLOT);
        const uint64_t m = Ops::load(slot + N);
        if (meta_status(m) == ST_EMPTY) return SN_NONE;
        bool same = true;
#pragma unroll
        for (int w = 0; w < N; w++) same &= (Ops::load(slot + w) == key[w]);
        if (same) { meta_out = m; return (uint64_t)b * t.W + s; }
      }
    }
    return SN_NONE;
  }
  const Probe pr = probe_of_key<N>(key);
  const uint64_t rep = (uint64_t)pr.tag8 * 0x0101010101010101ull;
  for (uint32_t r = 0; r <= t.max_rehash; ++r) {
    const uint32_t b = lookup3_key<N>(key, r).c & t.bucket_mask;
    uint8_t *base = bucket_base(t, b);
    uint32_t line = mulhi32(pr.h, t.NL);
    for (uint32_t step = 0; step < t.NL; ++step) {
      uint8_t *lp = base + (uint64_t)line * LINE_BYTES;
      const uint64_t hdr = Ops::load(reinterpret_cast<uint64_t *>(lp));
      const uint64_t valid = header_valid_mask((line == t.NL - 1) ? t.last_slots : G);
      uint64_t cand = zero_bytes64(hdr ^ rep) & valid;
      while (cand) {
        const uint32_t i = first_byte_index(cand);
        cand &= cand - 1;
        uint64_t *slot = reinterpret_cast<uint64_t *>(lp + 8u + i * SLOT);
        bool same = true;
#pragma unroll
        for (int w = 0; w < N; w++) same &= (Ops::load(slot + w) == key[w]);
        if (same) { meta_out = Ops::load(slot + N); return (uint64_t)b * t.W + line * G + i; }
      }
      if (zero_bytes64(hdr) & valid) return SN_NONE;  // a claimable slot ends the probe
      line = (line + 1 == t.NL) ? 0u : line + 1;
    }
  }
  return SN_NONE;
}

// --------------------------------------------------------------- degrees
LASV_HD uint32_t sn_side(uint32_t edges, uint32_t o) { return (edges >> (4u * o)) & 15u; }
LASV_HD uint32_t popc4(uint32_t x) { return (x & 1u) + ((x >> 1) & 1u) + ((x >> 2) & 1u) + ((x >> 3) & 1u); }
LASV_HD char base_char(uint32_t c) { return (char)((0x54474341u >> (8u * c)) & 0xFFu); }  // "ACGT"
LASV_HD uint32_t low_bit4(uint32_t x) { return (x & 1u) ? 0u : (x & 2u) ? 1u : (x & 4u) ? 2u : 3u; }
LASV_HD bool sn_interior(uint32_t edges) { return popc4(edges & 15u) == 1u && popc4(edges >> 4) == 1u; }
// the side through which a non-interior node triggers a path: reverse if it has
// exactly one reverse edge (dB_graph_population.c:1481-1486 tries that walk
// first), else forward if exactly one forward edge (:1512), else 2 = none
LASV_HD uint32_t sn_trigger_side(uint32_t edges) {
  if (popc4(edges >> 4) == 1u) return 1u;
  if (popc4(edges & 15u) == 1u) return 0u;
  return 2u;
}

// One step of db_graph_get_next_node (dB_graph_population.c:156-206): from node
// `key` read in orientation o along nucleotide nuc.  Returns the slot of the next
// node (SN_NONE: the edge points at nothing -- the reference die()s), its key,
// meta word and orientation, and the nucleotide that leads back.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD uint64_t sn_step(const TableView &t, int k, const uint64_t (&key)[N], uint32_t o, uint32_t nuc,
                         uint64_t (&key2)[N], uint64_t &meta2, uint32_t &o2, uint32_t &back) {
  uint64_t km[N], rc[N];
  if (o) kmer_revcomp<N>(key, k, km);
  else {
#pragma unroll
    for (int w = 0; w < N; w++) km[w] = key[w];
  }
  back = 3u - kmer_first_base<N>(km, k);
  kmer_shift_append<N>(km, nuc, k);
  kmer_revcomp<N>(km, k, rc);
  o2 = kmer_less<N>(rc, km) ? 1u : 0u;
#pragma unroll
  for (int w = 0; w < N; w++) key2[w] = o2 ? rc[w] : km[w];
  return sn_locate<N, Ops>(t, key2, meta2);
}

// ------------------------------------------------------------------ records
// One path between two non-interior nodes (or one cycle of interior nodes).
struct SnPath {
  uint64_t s_q;    // start node
  uint64_t t_q;    // end node
  uint64_t min_q;  // lowest interior slot (SN_NONE: none); cycles: the lowest slot of the cycle
  uint32_t len;    // edges = bases after the first k-mer
  uint32_t bits;   // see below
};
// bits: start (orientation, first nucleotide), the twin's start at the other end
// (orientation, first nucleotide), orientation in which the walk from s crosses the
// node min_q, trigger flags (s / t is a node that prints this path when the
// traversal reaches it unvisited), trigger sides of s and t, cycle flag
constexpr uint32_t SNB_S_O = 1u << 0, SNB_T_O = 1u << 3, SNB_MIN_O = 1u << 6, SNB_TRIG_S = 1u << 7,
                   SNB_TRIG_T = 1u << 8, SNB_CYCLE = 1u << 13,
                   SNB_BOTH_WAYS = 1u << 14;  // a cycle that crosses every node in both orientations
constexpr int SNB_S_N = 1, SNB_T_N = 4, SNB_S_SIDE = 9, SNB_T_SIDE = 11;

struct SnCounters {
  unsigned long long n_nodes;        // occupied slots
  unsigned long long n_interior;     // nodes with one edge on each side
  unsigned long long n_starts;       // path starts = edges of non-interior nodes (both twins)
  unsigned long long n_paths;        // records written by the path scan
  unsigned long long n_uncovered;    // interior nodes no path went through (cycles)
  unsigned long long n_cycles;       // cycle records written
  unsigned long long n_dangling;     // walks that met an edge to a node that is not in the table
  unsigned long long n_overflow;     // records that did not fit (sizing bug)
};

// node q as seen by the counting pass: 0 = empty, else 1 | interior << 1 | starts << 2
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD uint32_t sn_classify(const TableView &t, uint64_t q) {
  const uint32_t bucket = (uint32_t)(q / t.W);
  const uint64_t m = Ops::load(slot_ptr<N>(t, bucket, (uint32_t)(q - (uint64_t)bucket * t.W)) + N);
  if (meta_status(m) == ST_EMPTY) return 0u;
  const uint32_t e = meta_edges(m);
  if (sn_interior(e)) return 1u | 2u;
  return 1u | ((popc4(e & 15u) + popc4(e >> 4)) << 2);
}

// Walk the path that leaves node (s_q, key, edges) in orientation o along nuc.
// Fills rec and returns true if this start keeps the record (its twin does not);
// marks the interior nodes it crosses in covered[].  max_steps bounds a walk
// through a corrupt table.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool sn_walk_path(const TableView &t, int k, uint64_t s_q, const uint64_t (&s_key)[N], uint32_t s_edges,
                          uint32_t o, uint32_t nuc, uint64_t max_steps, uint8_t *covered, SnPath &rec,
                          bool &dangling) {
  uint64_t key[N], key2[N];
#pragma unroll
  for (int w = 0; w < N; w++) key[w] = s_key[w];
  uint32_t co = o, cn = nuc, o2 = 0, back = 0, min_o = 0;
  uint64_t q2 = SN_NONE, m2 = 0, min_q = SN_NONE, len = 0;
  dangling = false;
  for (;;) {
    q2 = sn_step<N, Ops>(t, k, key, co, cn, key2, m2, o2, back);
    if (q2 == SN_NONE) { dangling = true; return false; }
    len++;
    const uint32_t e2 = meta_edges(m2);
    if (!sn_interior(e2) || len >= max_steps) break;
    if (covered) covered[q2] = 1;
    if (q2 < min_q) { min_q = q2; min_o = o2; }
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = key2[w];
    co = o2;
    cn = low_bit4(sn_side(e2, o2));
  }
  // the twin starts at (t, !o2) along `back`; the smaller (state, nucleotide) keeps the record
  const uint64_t mine = ((2 * s_q + o) << 2) | nuc, twin = ((2 * q2 + (o2 ^ 1u)) << 2) | back;
  if (mine > twin) return false;
  const uint32_t e2 = meta_edges(m2);
  const uint32_t s_side = sn_trigger_side(s_edges), t_side = sn_trigger_side(e2);
  rec.s_q = s_q;
  rec.t_q = q2;
  rec.min_q = min_q;
  rec.len = (uint32_t)len;
  rec.bits = (o ? SNB_S_O : 0u) | (nuc << SNB_S_N) | ((o2 ^ 1u) ? SNB_T_O : 0u) | (back << SNB_T_N) |
             (min_o ? SNB_MIN_O : 0u) | (s_side == o ? SNB_TRIG_S : 0u) | (t_side == (o2 ^ 1u) ? SNB_TRIG_T : 0u) |
             (s_side << SNB_S_SIDE) | (t_side << SNB_T_SIDE);
  return true;
}

// Interior node q that no path crossed: it lies on a cycle of interior nodes.
// Walk it forward from (q, forward) back to (q, forward); the lowest slot of
// the cycle keeps the record (the reference prints a cycle from the first of
// its nodes the traversal meets, read forward: dB_graph_population.c:1481-1508).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool sn_walk_cycle(const TableView &t, int k, uint64_t q, const uint64_t (&q_key)[N], uint32_t q_edges,
                           uint64_t max_steps, SnPath &rec, bool &dangling) {
  uint64_t key[N], key2[N];
#pragma unroll
  for (int w = 0; w < N; w++) key[w] = q_key[w];
  const uint32_t n0 = low_bit4(sn_side(q_edges, 0u));
  uint32_t co = 0, cn = n0, o2 = 0, back = 0;
  uint64_t q2 = SN_NONE, m2 = 0, len = 0;
  bool both_ways = false;
  dangling = false;
  for (;;) {
    q2 = sn_step<N, Ops>(t, k, key, co, cn, key2, m2, o2, back);
    if (q2 == SN_NONE) { dangling = true; return false; }
    len++;
    if (q2 < q) return false;                 // a lower slot of the same cycle keeps the record
    if ((q2 == q && o2 == 0u) || len >= max_steps) break;
    if (q2 == q) both_ways = true;            // back at q in the other orientation: half way round
    const uint32_t e2 = meta_edges(m2);
    if (!sn_interior(e2)) { dangling = true; return false; }  // cannot happen in a consistent graph
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = key2[w];
    co = o2;
    cn = low_bit4(sn_side(e2, o2));
  }
  rec.s_q = q;
  rec.t_q = q;
  rec.min_q = q;
  rec.len = (uint32_t)len;
  rec.bits = SNB_CYCLE | (n0 << SNB_S_N) | (both_ways ? SNB_BOTH_WAYS : 0u);
  return true;
}

// One trigger of the replay: at slot `pos` the traversal reaches a node that would
// print path `path` (role 0: its lowest interior node, or a cycle; 1: its start node;
// 2: its end node -- the node must still be unvisited for roles 1 and 2).  Events are
// processed in slot order; they carry what the replay needs of the path record.
struct SnEvent {
  uint64_t s_q, t_q;
  uint32_t len, bits;
  uint32_t path;
  uint32_t role;
};

LASV_HD int sn_path_events(const SnPath &p, uint64_t (&pos)[3], uint32_t (&role)[3]) {
  if (p.bits & SNB_CYCLE) {
    pos[0] = p.min_q;
    role[0] = 0u;
    return 1;
  }
  int n = 0;
  if (p.min_q != SN_NONE) { pos[n] = p.min_q; role[n++] = 0u; }
  if (p.bits & SNB_TRIG_S) { pos[n] = p.s_q; role[n++] = 1u; }
  // a path that is its own twin carries both flags for the same (node, side): one event
  const bool self_twin = p.s_q == p.t_q && ((p.bits & SNB_S_O) != 0) == ((p.bits & SNB_T_O) != 0);
  if ((p.bits & SNB_TRIG_T) && !self_twin) { pos[n] = p.t_q; role[n++] = 2u; }
  return n;
}

// What the replay hands back for every printed supernode: where to start reading,
// and where its record starts in the FASTA image.
struct SnPrint {
  uint64_t q;       // first node
  uint64_t offset;  // byte offset of the record (header line) in the FASTA image
  uint32_t len;     // edges; the sequence has k + len bases
  uint32_t on;      // orientation | first nucleotide << 1
};

// ">node_<i> length:<bases> INFO:KMER=<k>\n" (dB_graph_population.c:2761, :6167)
LASV_HD uint32_t dec_digits(uint64_t v) {
  uint32_t d = 1;
  while (v >= 10) { v /= 10; d++; }
  return d;
}
LASV_HD uint32_t put_dec(char *dst, uint64_t v) {
  const uint32_t d = dec_digits(v);
  for (uint32_t i = d; i-- > 0;) { dst[i] = (char)('0' + (int)(v % 10)); v /= 10; }
  return d;
}
LASV_HD uint32_t sn_header_bytes(uint64_t i, uint64_t bases, int k) {
  return 6u + dec_digits(i) + 8u + dec_digits(bases) + 11u + dec_digits((uint64_t)k) + 1u;
}
LASV_HD uint32_t sn_write_header(char *dst, uint64_t i, uint64_t bases, int k) {
  const char a[] = ">node_", b[] = " length:", c[] = " INFO:KMER=";
  uint32_t n = 0;
  for (int j = 0; j < 6; j++) dst[n++] = a[j];
  n += put_dec(dst + n, i);
  for (int j = 0; j < 8; j++) dst[n++] = b[j];
  n += put_dec(dst + n, bases);
  for (int j = 0; j < 11; j++) dst[n++] = c[j];
  n += put_dec(dst + n, (uint64_t)k);
  dst[n++] = '\n';
  return n;
}

// Sequential byte output with 8-byte stores: single bytes up to the first 8-byte boundary,
// then whole words (one thread writes one record; neighbouring records never share a word store).
struct ByteWriter {
  char *p;       // next byte to be stored
  uint64_t acc;  // bytes collected for the word at p (p is 8-byte aligned whenever n > 0)
  uint32_t n;
  LASV_HD explicit ByteWriter(char *dst) : p(dst), acc(0), n(0) {}
  LASV_HD void put(char c) {
    if (n == 0 && (reinterpret_cast<uintptr_t>(p) & 7u) != 0) {
      *p++ = c;
      return;
    }
    acc |= (uint64_t)(uint8_t)c << (8u * n);
    if (++n == 8u) {
      *reinterpret_cast<uint64_t *>(p) = acc;  // little-endian: byte j of the word is p[j]
      p += 8;
      acc = 0;
      n = 0;
    }
  }
  LASV_HD void flush() {
    for (uint32_t j = 0; j < n; j++) p[j] = (char)(acc >> (8u * j));
    p += n;
    acc = 0;
    n = 0;
  }
};

// Record i of the FASTA: header line, then the first k-mer as read in the start
// orientation and one base per edge, then the newline
// (print_ultra_minimal_fasta_from_path).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool sn_write_record(const TableView &t, int k, const SnPrint &p, uint64_t i, char *image) {
  uint64_t key[N], key2[N], km[N];
  const uint64_t m = sn_load_node<N, Ops>(t, p.q, key);
  if (!m) return false;
  uint32_t co = p.on & 1u, cn = p.on >> 1, o2 = 0, back = 0;
  if (co) kmer_revcomp<N>(key, k, km);
  else {
#pragma unroll
    for (int w = 0; w < N; w++) km[w] = key[w];
  }
  char hdr[72];
  const uint32_t hn = sn_write_header(hdr, i, (uint64_t)k + p.len, k);
  ByteWriter out(image + p.offset);
  for (uint32_t b = 0; b < hn; b++) out.put(hdr[b]);
  for (int b = 0; b < k; b++) out.put(base_char(kmer_base_at<N>(km, k, b)));
  bool ok = true;
  for (uint32_t e = 0; e < p.len; e++) {
    uint64_t m2 = 0;
    out.put(base_char(cn));
    if (e + 1 == p.len) break;
    if (sn_step<N, Ops>(t, k, key, co, cn, key2, m2, o2, back) == SN_NONE) { ok = false; break; }
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = key2[w];
    co = o2;
    cn = low_bit4(sn_side(meta_edges(m2), o2));
  }
  out.put('\n');
  out.flush();
  return ok;
}

#if defined(__CUDACC__)
// launch interface (supernode_kernels.cu); c is a device SnCounters
void launch_sn_count(int N, const TableView &t, SnCounters *c, cudaStream_t st);
void launch_sn_starts(int N, const TableView &t, uint64_t *starts, uint64_t cap, unsigned long long *n_out,
                      SnCounters *c, cudaStream_t st);
void launch_sn_paths(int N, const TableView &t, int k, const uint64_t *starts, uint64_t n_starts,
                     uint8_t *covered, SnPath *recs, uint64_t cap, SnCounters *c, cudaStream_t st);
void launch_sn_uncovered(int N, const TableView &t, const uint8_t *covered, uint64_t *list, uint64_t cap,
                         SnCounters *c, cudaStream_t st);
void launch_sn_cycles(int N, const TableView &t, int k, const uint64_t *list, uint64_t n, SnPath *recs,
                      uint64_t cap, SnCounters *c, cudaStream_t st);
void launch_sn_write(int N, const TableView &t, int k, const SnPrint *prints, uint64_t n, char *image,
                     SnCounters *c, cudaStream_t st);
// events of recs[0..n) appended to (keys, vals) at *n_events; val = (index_base + i) << 2 | role
void launch_sn_events(const SnPath *recs, uint64_t n, uint32_t index_base, uint64_t *keys, uint32_t *vals,
                      unsigned long long *n_events, cudaStream_t st);
// sort the events by slot (keys_alt / vals_alt / tmp: scratch), then expand them to SnEvent in that order
size_t sn_sort_tmp_bytes(uint64_t n_events);
void launch_sn_sort_events(uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt, uint64_t n_events,
                           int key_bits, void *tmp, size_t tmp_bytes, const SnPath *paths, uint64_t n_paths,
                           const SnPath *cycles, SnEvent *events, cudaStream_t st);
#endif

}  // namespace lasv
