/*
 * oracle/dbg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of laSV's de Bruijn graph build
 * (FASTQ reads -> quality/N filter -> 2-bit k-mers -> canonical key ->
 * find-or-insert into the CORTEX bucketed hash with coverage + edge bits).
 * It follows the reference's own control flow function by function (each
 * function below names the reference file:line it restates) rather than the
 * closed-form formulation the CUDA kernels use, so the two are independent
 * derivations of the same result.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library, and only as the checker.  The product (lasv_b200/) never
 * links, imports or calls it.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle_pinned.py)
 * against (a) the known-answer vectors of SURVEY.md section 8c, (b) golden
 * .ctx fixtures produced by the compiled, unmodified reference
 * (oracle/_ref/dB_graph_{31,63,95}; generator: tests/golden/make_golden.py)
 * and (c) function-level calls into oracle/_ref/libref_*.so when present.
 *
 * N (number of 64-bit words per k-mer) is a run-time value here (1..3); the
 * reference fixes it at compile time (Makefile:8-43, MAXK).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_MAXN 3

typedef struct {
  uint64_t kmer[ORC_MAXN]; /* word 0 most significant (binary_kmer.h:41-47) */
  uint32_t covg;           /* element.h:77-82 */
  uint8_t edges;
  uint8_t status;          /* 0 = unassigned (empty), 1 = none */
} orc_node;

typedef struct {
  int k, N, L, W, max_rehash;
  uint64_t n_buckets;
  orc_node *table;
  uint64_t unique_kmers;
  uint64_t collisions[32];
  int table_full; /* set where the reference die()s (hash_table.c:309-317) */
} orc_graph;

typedef struct {
  uint64_t n_reads;
  uint64_t bad_reads;     /* reads with no k clean bases (file_reader.c:579,675,693) */
  uint64_t bases_read;    /* file_reader.c:584,768 */
  uint64_t bases_loaded;  /* file_reader.c:460 */
  uint64_t kmers_inserted;/* number of find_or_insert calls made by the loader */
} orc_stats;

/* ------------------------------------------------------------------------ */
/* event_encoding.c:13-36  char_to_binary_nucleotide */
static int orc_char_to_nuc(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4; /* Undefined */
  }
}

/* binary_kmer.c:115-131  binary_kmer_left_shift_one_base */
static void orc_left_shift_one_base(uint64_t *kmer, int k, int N) {
  int top_word = N - (k + 31) / 32;
  int i;
  for (i = top_word; i < N - 1; i++) {
    kmer[i] <<= 2;
    kmer[i] |= (kmer[i + 1] >> 62);
  }
  kmer[N - 1] <<= 2;
  int top_bits = 2 * (k % 32);
  kmer[top_word] &= (~(uint64_t)0 >> (64 - top_bits));
}

/* binary_kmer.c:134-143 */
static void orc_shift_insert(uint64_t *kmer, int nuc, int k, int N) {
  orc_left_shift_one_base(kmer, k, N);
  kmer[N - 1] |= (uint64_t)nuc;
}

/* binary_kmer.c:388-417  seq_to_binary_kmer.  Returns 0 on undefined char. */
int orc_seq_to_kmer(const char *seq, int k, int N, uint64_t *out) {
  int j;
  for (j = 0; j < N; j++) out[j] = 0;
  for (j = 0; j < k; j++) {
    int n = orc_char_to_nuc(seq[j]);
    if (n == 4) return 0;
    orc_shift_insert(out, n, k, N);
  }
  return 1;
}

/* binary_kmer.c:428-452  binary_kmer_to_seq */
void orc_kmer_to_seq(const uint64_t *kmer, int k, int N, char *seq) {
  uint64_t tmp[ORC_MAXN];
  int i, j;
  for (i = 0; i < N; i++) tmp[i] = kmer[i];
  for (j = k - 1; j >= 0; j--) {
    seq[j] = "ACGT"[tmp[N - 1] & 3];
    /* binary_kmer.c:100-112 right shift one base */
    for (i = N - 1; i > 0; i--) {
      tmp[i] >>= 2;
      tmp[i] |= (tmp[i - 1] << 62);
    }
    tmp[0] >>= 2;
  }
  seq[k] = '\0';
}

/* binary_kmer.c:456-523  binary_kmer_reverse_complement: complement every
 * word, then pull bases off the right end of the source and push them on the
 * right end of the destination, one base at a time. */
void orc_revcomp(const uint64_t *kmer, int k, int N, uint64_t *out) {
  uint64_t copy[ORC_MAXN];
  int i, j, w;
  for (i = 0; i < N; i++) { copy[i] = ~kmer[i]; out[i] = 0; }
  for (i = N - 1; i > 0; i--) {
    for (j = 0; j < 32; j++) {
      for (w = 0; w < N - 1; w++) {
        out[w] <<= 2;
        out[w] |= (out[w + 1] >> 62);
      }
      out[N - 1] <<= 2;
      out[N - 1] |= copy[i] & 3;
      copy[i] >>= 2;
    }
  }
  int top_bases = k % 32;
  for (i = 0; i < top_bases; i++) {
    for (w = 0; w < N - 1; w++) {
      out[w] <<= 2;
      out[w] |= (out[w + 1] >> 62);
    }
    out[N - 1] <<= 2;
    out[N - 1] |= copy[0] & 3;
    copy[0] >>= 2;
  }
  out[0] &= (~(uint64_t)0 >> (64 - 2 * top_bases));
}

/* binary_kmer.c:66-97  binary_kmer_less_than */
static int orc_less_than(const uint64_t *l, const uint64_t *r, int k, int N) {
  int i;
  for (i = N - k / 32 - 1; i < N; i++) {
    if (l[i] < r[i]) return 1;
    if (l[i] > r[i]) return 0;
  }
  return 0;
}

static int orc_kmer_eq(const uint64_t *a, const uint64_t *b, int N) {
  int i;
  for (i = 0; i < N; i++) if (a[i] != b[i]) return 0;
  return 1;
}

/* element.c:308-336  element_get_key */
void orc_get_key(const uint64_t *kmer, int k, int N, uint64_t *key) {
  int top_bits = 2 * (k % 32);
  int first = (int)((kmer[0] >> (top_bits - 2)) & 3); /* binary_kmer.c:525 */
  int last = (int)(kmer[N - 1] & 3);                  /* binary_kmer.c:532 */
  int rev_last = ~last & 3;
  int i;
  if (first < rev_last) {
    for (i = 0; i < N; i++) key[i] = kmer[i];
  } else if (first > rev_last) {
    orc_revcomp(kmer, k, N, key);
  } else {
    orc_revcomp(kmer, k, N, key);
    if (orc_less_than(kmer, key, k, N))
      for (i = 0; i < N; i++) key[i] = kmer[i];
  }
}

/* hash_value.c:114-122, 149-158, 282-449  lookup3 hashlittle, byte-wise
 * (the reference takes the aligned 32-bit path; both paths are defined to
 * return the same value on little-endian machines). */
#define ORC_ROT(x, k) (((x) << (k)) | ((x) >> (32 - (k))))
uint32_t orc_hashlittle(const uint8_t *k, size_t length, uint32_t initval) {
  uint32_t a, b, c;
  a = b = c = 0xdeadbeef + ((uint32_t)length) + initval;
  while (length > 12) {
    a += k[0] + ((uint32_t)k[1] << 8) + ((uint32_t)k[2] << 16) + ((uint32_t)k[3] << 24);
    b += k[4] + ((uint32_t)k[5] << 8) + ((uint32_t)k[6] << 16) + ((uint32_t)k[7] << 24);
    c += k[8] + ((uint32_t)k[9] << 8) + ((uint32_t)k[10] << 16) + ((uint32_t)k[11] << 24);
    a -= c; a ^= ORC_ROT(c, 4);  c += b;
    b -= a; b ^= ORC_ROT(a, 6);  a += c;
    c -= b; c ^= ORC_ROT(b, 8);  b += a;
    a -= c; a ^= ORC_ROT(c, 16); c += b;
    b -= a; b ^= ORC_ROT(a, 19); a += c;
    c -= b; c ^= ORC_ROT(b, 4);  b += a;
    length -= 12;
    k += 12;
  }
  switch (length) {
    case 12: c += ((uint32_t)k[11]) << 24; /* fall through */
    case 11: c += ((uint32_t)k[10]) << 16; /* fall through */
    case 10: c += ((uint32_t)k[9]) << 8;   /* fall through */
    case 9:  c += k[8];                    /* fall through */
    case 8:  b += ((uint32_t)k[7]) << 24;  /* fall through */
    case 7:  b += ((uint32_t)k[6]) << 16;  /* fall through */
    case 6:  b += ((uint32_t)k[5]) << 8;   /* fall through */
    case 5:  b += k[4];                    /* fall through */
    case 4:  a += ((uint32_t)k[3]) << 24;  /* fall through */
    case 3:  a += ((uint32_t)k[2]) << 16;  /* fall through */
    case 2:  a += ((uint32_t)k[1]) << 8;   /* fall through */
    case 1:  a += k[0]; break;
    case 0:  return c;
  }
  c ^= b; c -= ORC_ROT(b, 14);
  a ^= c; a -= ORC_ROT(c, 11);
  b ^= a; b -= ORC_ROT(a, 25);
  c ^= b; c -= ORC_ROT(b, 16);
  a ^= c; a -= ORC_ROT(c, 4);
  b ^= a; b -= ORC_ROT(a, 14);
  c ^= b; c -= ORC_ROT(b, 24);
  return c;
}

/* hash_value.c:767-773 hash_value + hash_table.c:73-78 (rehash added to the
 * last word before hashing).  Returns the bucket index for 2^L buckets. */
uint32_t orc_hash_value(const uint64_t *key, int N, int rehash, int L) {
  uint64_t tmp[ORC_MAXN];
  int i;
  for (i = 0; i < N; i++) tmp[i] = key[i];
  tmp[N - 1] += (uint64_t)rehash;
  int hashval = (int)orc_hashlittle((const uint8_t *)tmp, 8 * (size_t)N, 10);
  int number_buckets = (int)((long long)1 << L);
  return (uint32_t)(hashval & (number_buckets - 1));
}

/* ------------------------------------------------------------------------ */
/* hash_table.c:13-53  hash_table_new */
orc_graph *orc_graph_new(int k, int L, int W, int max_rehash) {
  orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
  if (!g) return NULL;
  g->k = k;
  g->N = (2 * k) / 64 + 1; /* graph_operations.c:620 */
  g->L = L;
  g->W = W;
  g->max_rehash = max_rehash;
  g->n_buckets = (uint64_t)1 << L;
  g->table = (orc_node *)calloc(g->n_buckets * (uint64_t)W, sizeof(orc_node));
  if (!g->table || g->N > ORC_MAXN) { free(g->table); free(g); return NULL; }
  return g;
}

void orc_graph_free(orc_graph *g) {
  if (g) { free(g->table); free(g); }
}

/* hash_table.c:69-122  hash_table_find_in_bucket */
static int orc_find_in_bucket(orc_graph *g, const uint64_t *key, uint64_t *pos,
                              int *overflow, int rehash) {
  uint32_t hv = orc_hash_value(key, g->N, rehash, g->L);
  int found = 0, i = 0;
  *overflow = 0;
  *pos = (uint64_t)hv * (uint64_t)g->W;
  while (i < g->W && g->table[*pos].status != 0 && !found) {
    if (orc_kmer_eq(key, g->table[*pos].kmer, g->N)) found = 1;
    else { (*pos)++; i++; }
  }
  if (i == g->W) *overflow = 1;
  return found;
}

/* hash_table.c:270-327  hash_table_find_or_insert (+ element_initialise,
 * element.c:339-362).  Returns NULL where the reference dies (table full). */
static orc_node *orc_find_or_insert(orc_graph *g, const uint64_t *key, int *found) {
  int rehash = 0, overflow, i;
  uint64_t pos;
  orc_node *ret = NULL;
  do {
    *found = orc_find_in_bucket(g, key, &pos, &overflow, rehash);
    if (!*found) {
      if (!overflow) {
        orc_node *e = &g->table[pos];
        for (i = 0; i < g->N; i++) e->kmer[i] = key[i];
        e->covg = 0;
        e->edges = 0;
        e->status = 1; /* none */
        g->unique_kmers++;
        ret = e;
      } else {
        rehash++;
        if (rehash > g->max_rehash) { g->table_full = 1; return NULL; }
      }
    } else {
      ret = &g->table[pos];
    }
  } while (overflow);
  g->collisions[rehash]++;
  return ret;
}

/* hash_table.c:234-267  hash_table_find */
static orc_node *orc_find(orc_graph *g, const uint64_t *key) {
  int rehash = 0, overflow, found;
  uint64_t pos;
  do {
    found = orc_find_in_bucket(g, key, &pos, &overflow, rehash);
    if (found) return &g->table[pos];
    if (overflow) {
      rehash++;
      if (rehash > g->max_rehash) return NULL;
    }
  } while (overflow);
  return NULL;
}

/* element.c:470-492  db_node_get_orientation: 0 forward, 1 reverse */
static int orc_get_orientation(const uint64_t *kmer, const orc_node *e, int k, int N) {
  if (orc_kmer_eq(e->kmer, kmer, N)) return 0;
  uint64_t rc[ORC_MAXN];
  orc_revcomp(kmer, k, N, rc);
  if (orc_kmer_eq(e->kmer, rc, N)) return 1;
  fprintf(stderr, "oracle: orientation programming error\n");
  abort();
}

/* element.c:501-513  db_node_add_labeled_edge */
static void orc_add_labeled_edge(orc_node *e, int orient, int base) {
  uint8_t edge = (uint8_t)(1u << base);
  if (orient == 1) edge <<= 4;
  e->edges |= edge; /* add_edges, element.c:223-228 */
}

/* element.c:519-560  db_node_add_edge */
static void orc_add_edge(orc_node *src, orc_node *tgt, int src_o, int tgt_o, int k, int N) {
  uint64_t src_k[ORC_MAXN], tgt_k[ORC_MAXN], tmp[ORC_MAXN];
  int i;
  for (i = 0; i < N; i++) { src_k[i] = src->kmer[i]; tgt_k[i] = tgt->kmer[i]; }
  if (src_o == 1) { orc_revcomp(src_k, k, N, tmp); for (i = 0; i < N; i++) src_k[i] = tmp[i]; }
  if (tgt_o == 1) { orc_revcomp(tgt_k, k, N, tmp); for (i = 0; i < N; i++) tgt_k[i] = tmp[i]; }
  orc_add_labeled_edge(src, src_o, (int)(tgt_k[N - 1] & 3));
  orc_revcomp(src_k, k, N, tmp);
  orc_add_labeled_edge(tgt, tgt_o ^ 1, (int)(tmp[N - 1] & 3));
}

/* element.c:395-417  db_node_update_coverage (saturating) */
static void orc_update_coverage(orc_node *e, uint32_t update) {
  if (0xFFFFFFFFu - update >= e->covg) e->covg += update;
  else e->covg = 0xFFFFFFFFu;
}

/* ------------------------------------------------------------------------ */
/* In-memory stand-in for the SeqFile cursor (libs/seq_file/seq_file.c:700-840,
 * seq_fastq.c:117-150): bases are buffered per read, qualities advance in
 * lock-step; reading past the end fails and ends the read. */
typedef struct {
  const char *bases, *quals; /* quals may be NULL (FASTA: no quality filter) */
  uint64_t len, off;
  int entry_read;
} orc_cursor;

#define ORC_IS_BASE(x) ((x)=='a'||(x)=='A'||(x)=='c'||(x)=='C'||(x)=='g'||(x)=='G'||(x)=='t'||(x)=='T')

static char orc_upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

/* file_reader.c:180-207  _read_base */
static int orc_read_base(orc_cursor *c, char *b, char *q) {
  if (!c->entry_read || c->off >= c->len) { c->entry_read = 0; return 0; }
  *b = c->bases[c->off];
  if (c->quals) *q = c->quals[c->off];
  c->off++;
  if (!ORC_IS_BASE(*b) && *b != 'n' && *b != 'N') *b = 'N';
  *b = orc_upper(*b);
  return 1;
}

/* file_reader.c:209-230  _read_k_bases (seq_file.c:795-840: a short read
 * consumes what is left and fails) */
static int orc_read_k_bases(orc_cursor *c, char *bases, char *quals, int k) {
  int i;
  if (!c->entry_read) return 0;
  for (i = 0; i < k; i++) {
    if (c->off >= c->len) { c->entry_read = 0; return 0; }
    bases[i] = orc_upper(c->bases[c->off]);
    if (c->quals) quals[i] = c->quals[c->off];
    c->off++;
  }
  return 1;
}

/* file_reader.c:120-178  _kmer_errors with homopolymer_cutoff == 0 (laSV never
 * sets it: graph_operations.c:741-742).  Note the quality loop OVERWRITES the
 * skip found by the base loop, exactly as the reference does. */
static int orc_kmer_errors(const char *kmer_str, const char *qual_str, int k,
                           int read_qual, signed char cutoff) {
  int skip = 0, i;
  for (i = k - 1; i >= 0; i--) {
    if (!ORC_IS_BASE(kmer_str[i])) { skip = i + 1; break; }
  }
  if (read_qual) {
    for (i = k - 1; i >= 0; i--) {
      if ((signed char)qual_str[i] <= cutoff) { skip = i + 1; break; }
    }
  }
  return skip;
}

/* file_reader.c:259-320  _read_first_kmer (first_base == 0 form) */
static int orc_read_first_kmer(orc_cursor *c, char *kmer_str, char *qual_str, int k,
                               int read_qual, signed char cutoff) {
  int skip;
  if (!orc_read_k_bases(c, kmer_str, qual_str, k)) return 0;
  while ((skip = orc_kmer_errors(kmer_str, qual_str, k, read_qual, cutoff)) > 0) {
    while (skip == k) {
      if (!orc_read_k_bases(c, kmer_str, qual_str, k)) return 0;
      skip = 0; /* skip_homorun_base == 0: the inner scan matches nothing */
    }
    if (skip > 0) {
      int keep = k - skip;
      memmove(kmer_str, kmer_str + skip, (size_t)keep);
      memmove(qual_str, qual_str + skip, (size_t)keep);
      if (!orc_read_k_bases(c, kmer_str + keep, qual_str + keep, skip)) return 0;
    }
  }
  return 1;
}

/* file_reader.c:322-473  _process_read */
static int orc_process_read(orc_graph *g, orc_cursor *c, char *kmer_str, char *qual_str,
                            int read_qual, signed char cutoff, uint64_t *curr_kmer,
                            orc_node *curr_node, int curr_orient, orc_stats *st,
                            uint64_t *hist, uint64_t hist_size) {
  int k = g->k, N = g->N, found;
  orc_node *prev_node = NULL;
  int prev_orient = 0;
  uint64_t key[ORC_MAXN];
  char base = 0, qual = 0;
  int keep_reading = 1, is_first = 1;
  while (keep_reading) {
    if (!is_first) {
      orc_seq_to_kmer(kmer_str, k, N, curr_kmer);
      orc_get_key(curr_kmer, k, N, key);
      curr_node = orc_find_or_insert(g, key, &found);
      if (!curr_node) return 0;
      st->kmers_inserted++;
      curr_orient = orc_get_orientation(curr_kmer, curr_node, k, N);
      orc_update_coverage(curr_node, 1);
    }
    is_first = 0;
    uint64_t contig_length = (uint64_t)k;
    while ((keep_reading = orc_read_base(c, &base, &qual))) {
      if (base == 'N' || (read_qual && (signed char)qual <= cutoff)) {
        keep_reading = orc_read_first_kmer(c, kmer_str, qual_str, k, read_qual, cutoff);
        break;
      }
      contig_length++;
      prev_node = curr_node;
      prev_orient = curr_orient;
      orc_shift_insert(curr_kmer, orc_char_to_nuc(base), k, N);
      orc_get_key(curr_kmer, k, N, key);
      curr_node = orc_find_or_insert(g, key, &found);
      if (!curr_node) return 0;
      st->kmers_inserted++;
      curr_orient = orc_get_orientation(curr_kmer, curr_node, k, N);
      orc_update_coverage(curr_node, 1);
      orc_add_edge(prev_node, curr_node, prev_orient, curr_orient, k, N);
    }
    st->bases_loaded += contig_length;
    if (hist) {
      uint64_t b = contig_length < hist_size - 1 ? contig_length : hist_size - 1;
      hist[b]++;
    }
  }
  return 1;
}

/* file_reader.c:475-587 load_se_seq_data_into_graph_colour (paired == 0) and
 * file_reader.c:589-773 load_pe_seq_data_into_graph_colour (paired != 0), with
 * duplicate removal off (laSV never enables it).  The paired loader looks up the
 * first k-mer of BOTH mates (:654-694) before it walks mate 1 and then mate 2
 * (:737-764); that only changes the order in which slots are claimed, never the
 * node multiset, and is kept so the oracle's table layout equals the reference's.
 *
 * Reads are given in memory: read i occupies seq[offsets[i] .. offsets[i+1]-sep)
 * (and the same span of qual when qual != NULL); `sep` separator bytes follow
 * each read; in paired mode reads 2p and 2p+1 are the mates of pair p.
 * `cutoff` is the RAW ASCII cutoff, i.e. threshold + offset
 * (file_reader.c:798,919), compared as signed char (file_reader.c:150,389).
 * Returns 1, or 0 if the table filled up (the reference would die()).       */
int orc_load_reads(orc_graph *g, const char *seq, const char *qual,
                   const uint64_t *offsets, uint64_t n_reads, int sep, int cutoff, int paired,
                   orc_stats *st, uint64_t *hist, uint64_t hist_size) {
  int k = g->k, N = g->N, found;
  int group = paired ? 2 : 1, m;
  char *kmer_str[2], *qual_str[2];
  uint64_t curr_kmer[2][ORC_MAXN], key[ORC_MAXN];
  orc_node *node[2];
  orc_cursor cur[2];
  int got[2], orient[2];
  uint64_t r;
  int read_qual = qual != NULL, ok = 1;
  for (m = 0; m < 2; m++) {
    kmer_str[m] = (char *)malloc((size_t)k + 1);
    qual_str[m] = (char *)calloc((size_t)k + 1, 1);
  }
  for (r = 0; r + (uint64_t)group <= n_reads && ok; r += (uint64_t)group) {
    for (m = 0; m < group; m++) {
      orc_cursor *c = &cur[m];
      c->bases = seq + offsets[r + m];
      c->quals = qual ? qual + offsets[r + m] : NULL;
      c->len = offsets[r + m + 1] - offsets[r + m] - (uint64_t)sep;
      c->off = 0;
      c->entry_read = 1;
      st->n_reads++;
      st->bases_read += c->len;
    }
    for (m = 0; m < group; m++)
      got[m] = orc_read_first_kmer(&cur[m], kmer_str[m], qual_str[m], k, read_qual, (signed char)cutoff);
    for (m = 0; m < group && ok; m++) {
      if (got[m]) {
        orc_seq_to_kmer(kmer_str[m], k, N, curr_kmer[m]);
        orc_get_key(curr_kmer[m], k, N, key);
        node[m] = orc_find_or_insert(g, key, &found);
        if (!node[m]) { ok = 0; break; }
        st->kmers_inserted++;
        orient[m] = orc_get_orientation(curr_kmer[m], node[m], k, N);
      } else {
        st->bad_reads++;
      }
    }
    for (m = 0; m < group && ok; m++) {
      if (!got[m]) continue;
      orc_update_coverage(node[m], 1);
      ok = orc_process_read(g, &cur[m], kmer_str[m], qual_str[m], read_qual, (signed char)cutoff,
                            curr_kmer[m], node[m], orient[m], st, hist, hist_size);
    }
  }
  for (m = 0; m < 2; m++) { free(kmer_str[m]); free(qual_str[m]); }
  return ok;
}

/* ------------------------------------------------------------------------ */
uint64_t orc_unique_kmers(const orc_graph *g) { return g->unique_kmers; }
int orc_table_full(const orc_graph *g) { return g->table_full; }
void orc_collisions(const orc_graph *g, uint64_t *out16) {
  int i;
  for (i = 0; i < 16; i++) out16[i] = g->collisions[i];
}

/* hash_table.c:178-187 traverse + element.c:894-910 record print: records in
 * TABLE ORDER, each kmer[N] (8N bytes) + uint32 covg + uint8 edges.
 * Returns number of records written (buffer must hold unique_kmers records). */
uint64_t orc_dump_records(const orc_graph *g, uint8_t *out) {
  uint64_t i, n = 0, cap = g->n_buckets * (uint64_t)g->W;
  size_t rec = 8 * (size_t)g->N + 5;
  for (i = 0; i < cap; i++) {
    const orc_node *e = &g->table[i];
    if (e->status != 1 && e->status != 2) continue; /* db_node_check_status_not_pruned, element.c:795-802 */
    memcpy(out + n * rec, e->kmer, 8 * (size_t)g->N);
    memcpy(out + n * rec + 8 * (size_t)g->N, &e->covg, 4);
    out[n * rec + 8 * (size_t)g->N + 4] = e->edges;
    n++;
  }
  return n;
}

/* test helper: status byte of every dumped record, in the order of orc_dump_records */
uint64_t orc_dump_status(const orc_graph *g, uint8_t *out) {
  uint64_t i, n = 0, cap = g->n_buckets * (uint64_t)g->W;
  for (i = 0; i < cap; i++) {
    const orc_node *e = &g->table[i];
    if (e->status != 1 && e->status != 2) continue;
    out[n++] = e->status;
  }
  return n;
}

/* file_reader.c:1686-1703 + hash_table.c:333-382: reload records through
 * find_or_insert semantics (keys are unique in a .ctx), adding covg/edges. */
int orc_load_records(orc_graph *g, const uint8_t *recs, uint64_t n) {
  size_t rec = 8 * (size_t)g->N + 5;
  uint64_t i, key[ORC_MAXN];
  int found;
  for (i = 0; i < n; i++) {
    uint32_t covg;
    memcpy(key, recs + i * rec, 8 * (size_t)g->N);
    memcpy(&covg, recs + i * rec + 8 * (size_t)g->N, 4);
    orc_node *e = orc_find_or_insert(g, key, &found);
    if (!e) return 0;
    orc_update_coverage(e, covg);
    e->edges |= recs[i * rec + 8 * (size_t)g->N + 4];
  }
  return 1;
}

/* Look a key up (hash_table_find); returns 1 and fills covg/edges if present. */
int orc_lookup(orc_graph *g, const uint64_t *key, uint32_t *covg, uint8_t *edges) {
  orc_node *e = orc_find(g, key);
  if (!e) return 0;
  *covg = e->covg;
  *edges = e->edges;
  return 1;
}

/* file_reader.c:2930-2994  print_binary_signature_NEW for one colour with the
 * default GraphInfo (graph_info.c: sample id "undefined", seq_err 0.01L, no
 * cleaning): writes the .ctx v6 header into out (>= 128 bytes), returns its
 * length (94 with these defaults). */
int orc_ctx_header(int k, int N, uint32_t mean_read_len, uint64_t total_seq, uint8_t *out) {
  uint8_t *p = out;
  int32_t v;
  memcpy(p, "CORTEX", 6); p += 6;
  v = 6; memcpy(p, &v, 4); p += 4;          /* BINVERSION */
  v = k; memcpy(p, &v, 4); p += 4;
  v = N; memcpy(p, &v, 4); p += 4;
  v = 1; memcpy(p, &v, 4); p += 4;          /* number of colours */
  v = (int32_t)mean_read_len; memcpy(p, &v, 4); p += 4;
  { long long t = (long long)total_seq; memcpy(p, &t, 8); p += 8; }
  v = 9; memcpy(p, &v, 4); p += 4; memcpy(p, "undefined", 9); p += 9;
  { long double e = (long double)0.01; /* graph_info.c:127 assigns the DOUBLE literal */ memset(p, 0, 16); memcpy(p, &e, sizeof(long double)); p += 16; }
  memset(p, 0, 4); p += 4;                  /* 4 x boolean (1 byte each) cleaning flags */
  v = -1; memcpy(p, &v, 4); p += 4;
  v = -1; memcpy(p, &v, 4); p += 4;
  v = 9; memcpy(p, &v, 4); p += 4; memcpy(p, "undefined", 9); p += 9;
  memcpy(p, "CORTEX", 6); p += 6;
  return (int)(p - out);
}

/* ------------------------------------------------------------------------ */
/* Supernode (maximal unambiguous contig) printing, restated step by step from
 * the reference: the sequential table traversal with `visited` marks, the
 * reverse walk to the end of the supernode, the forward walk that prints it.
 * The product (lasv_b200/csrc/supernodes.cuh) derives the same output from a
 * parallel path decomposition plus an event replay; this is its checker. */

/* element.c:653-679  db_node_has_precisely_one_edge_in_subgraph_...; *nuc is
 * left at the LAST edge seen, as in the reference */
static int orc_has_precisely_one_edge(const orc_node *e, int orient, int *nuc) {
  unsigned edges = e->edges;
  int n, count = 0;
  if (orient == 1) edges >>= 4;
  for (n = 0; n < 4; n++) {
    if (edges & 1u) { *nuc = n; count++; }
    edges >>= 1;
  }
  return count == 1;
}

/* dB_graph_population.c:156-206  db_graph_get_next_node_in_subgraph_...
 * (status pruned never occurs here; a node without edges is "not in the
 * subgraph", element.c:850-868) */
static orc_node *orc_get_next_node(orc_graph *g, const orc_node *cur, int cur_o, int *next_o,
                                   int edge, int *reverse_edge) {
  uint64_t local[ORC_MAXN], rev[ORC_MAXN], key[ORC_MAXN];
  int i;
  orc_node *next;
  for (i = 0; i < g->N; i++) local[i] = cur->kmer[i];
  orc_revcomp(local, g->k, g->N, rev);
  if (cur_o == 1) {
    *reverse_edge = (int)(local[g->N - 1] & 3);
    for (i = 0; i < g->N; i++) local[i] = rev[i];
  } else {
    *reverse_edge = (int)(rev[g->N - 1] & 3);
  }
  orc_shift_insert(local, edge, g->k, g->N);
  orc_get_key(local, g->k, g->N, key);
  next = orc_find(g, key);
  if (next) *next_o = orc_get_orientation(local, next, g->k, g->N);
  if (!next || next->status == 3 || next->edges == 0) return NULL; /* element.c:850-868 */
  return next;
}

typedef struct {
  orc_node **nodes;
  int *orients;
  int *labels;
  char *seq;
} orc_path;

/* dB_graph_population.c:783-896  ..._get_perfect_path_with_first_edge_...;
 * mark != 0: node_action = set status visited (2) on interior nodes.
 * Returns -1 where the reference die()s (edge to a node that is not there). */
static int orc_perfect_path_with_first_edge(orc_graph *g, orc_node *node, int orient, int limit,
                                            int fst_nuc, int mark, orc_path *p, int *is_cycle) {
  static const char nuc_char[4] = {'A', 'C', 'G', 'T'};
  orc_node *cur = node, *next = NULL;
  int cur_o = orient, next_o = 0, nuc = fst_nuc, rev_nuc, nuc2, length = 0;
  if (node->edges == 0) return 0; /* "This node is not in the graph of this person" */
  *is_cycle = 0;
  p->nodes[0] = node;
  p->orients[0] = orient;
  do {
    if (length > 0 && mark) cur->status = 2;
    next = orc_get_next_node(g, cur, cur_o, &next_o, nuc, &rev_nuc);
    if (!next) return -1;
    p->labels[length] = nuc;
    p->seq[length] = nuc_char[nuc];
    length++;
    p->nodes[length] = next;
    p->orients[length] = next_o;
    cur = next;
    cur_o = next_o;
  } while (length < limit && !(next == node && next_o == orient) &&
           orc_has_precisely_one_edge(next, next_o ^ 1, &nuc2) &&
           orc_has_precisely_one_edge(cur, cur_o, &nuc));
  if (next == node && next_o == orient) *is_cycle = 1;
  p->seq[length] = '\0';
  return length;
}

/* dB_graph_population.c:900-940  ..._get_perfect_path_... */
static int orc_perfect_path(orc_graph *g, orc_node *node, int orient, int limit, int mark,
                            orc_path *p, int *is_cycle) {
  int nuc;
  p->nodes[0] = node;
  p->orients[0] = orient;
  if (orc_has_precisely_one_edge(node, orient, &nuc))
    return orc_perfect_path_with_first_edge(g, node, orient, limit, nuc, mark, p, is_cycle);
  p->seq[0] = '\0';
  return 0;
}

/* dB_graph_population.c:1454-1527  db_graph_supernode_in_subgraph_... */
static int orc_supernode(orc_graph *g, orc_node *node, int limit, orc_path *p, int *is_cycle) {
  int is_cycler, length, length_reverse;
  length_reverse = orc_perfect_path(g, node, 1, limit, 0, p, &is_cycler);
  if (length_reverse < 0) return -1;
  if (length_reverse > 0) {
    int label, next_o;
    orc_node *lst = orc_get_next_node(g, p->nodes[length_reverse - 1], p->orients[length_reverse - 1],
                                      &next_o, p->labels[length_reverse - 1], &label);
    if (lst != p->nodes[length_reverse]) return -1; /* "db_graph_supernode broken!" */
    length = orc_perfect_path_with_first_edge(g, p->nodes[length_reverse],
                                              p->orients[length_reverse] ^ 1, limit, label, 1, p,
                                              is_cycle);
  } else {
    length = orc_perfect_path(g, node, 0, limit, 1, p, is_cycle);
  }
  if (length < 0) return -1;
  p->nodes[0]->status = 2;
  p->nodes[length]->status = 2;
  return length;
}

/* dB_graph_population.c:2696-2803  db_graph_print_supernodes_... with
 * filename_sings == "" (graph_operations.c:1281-1287), record format of
 * print_ultra_minimal_fasta_from_path (:6125-6178), then the status reset of
 * graph_operations.c:1290.  counts[0] = supernodes printed, [1] = singletons,
 * [2] = nodes visited, [3] = supernodes of exactly max_length edges.
 * Returns 0, or -1 on I/O error / inconsistent graph. */
int orc_print_supernodes(orc_graph *g, const char *path, int max_length, uint64_t *counts) {
  uint64_t i, cap = g->n_buckets * (uint64_t)g->W;
  orc_path p;
  int rc = 0;
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  p.nodes = (orc_node **)calloc((size_t)max_length + 1, sizeof(orc_node *));
  p.orients = (int *)calloc((size_t)max_length + 1, sizeof(int));
  p.labels = (int *)calloc((size_t)max_length + 1, sizeof(int));
  p.seq = (char *)calloc((size_t)max_length + 1 + (size_t)g->k, 1);
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  for (i = 0; i < cap && rc == 0; i++) {
    orc_node *e = &g->table[i];
    int is_cycle = 0, length;
    if (e->status == 0) continue; /* hash_table_traverse skips unassigned slots */
    counts[2]++;
    if (e->status != 1) continue;
    length = orc_supernode(g, e, max_length, &p, &is_cycle);
    if (length < 0) { rc = -1; break; }
    if (length > 0) {
      uint64_t fst[ORC_MAXN];
      char *fst_seq = (char *)malloc((size_t)g->k + 1);
      int w;
      if (p.orients[0] == 1) orc_revcomp(p.nodes[0]->kmer, g->k, g->N, fst);
      else for (w = 0; w < g->N; w++) fst[w] = p.nodes[0]->kmer[w];
      orc_kmer_to_seq(fst, g->k, g->N, fst_seq);
      fst_seq[g->k] = '\0';
      fprintf(f, ">node_%llu length:%i INFO:KMER=%d\n%s%s\n", (unsigned long long)counts[0],
              length + g->k, g->k, fst_seq, p.seq);
      free(fst_seq);
      if (length == max_length) counts[3]++;
      counts[0]++;
    } else {
      counts[1]++;
    }
  }
  for (i = 0; i < cap; i++)
    if (g->table[i].status == 2) g->table[i].status = 1;
  free(p.nodes); free(p.orients); free(p.labels); free(p.seq);
  if (fclose(f) != 0) rc = -1;
  return rc;
}

/* ------------------------------------------------------------------------ */
/* Cleaning, as `--remove_low_coverage_supernodes T` runs it (graph_operations.c:
 * 1069-1090): clip tips, then prune low-coverage supernodes.  Both mutate the
 * graph while they traverse it, so what they leave depends on the slot order. */

/* element.c:249-267  reset_one_edge */
static void orc_reset_one_edge(orc_node *e, int orient, int nuc) {
  e->edges &= (uint8_t)~(1u << (nuc + 4 * orient));
}

/* dB_graph_population.c:365-441  ..._clip_tip_with_orientation_...; node_action =
 * db_node_action_set_status_pruned, edges of the tip reset, back edge of the
 * junction reset.  Returns the tip length, 0 if nothing was clipped, -1 where the
 * reference die()s. */
static int orc_clip_tip_with_orientation(orc_graph *g, orc_node *node, int orientation, int limit,
                                         orc_node **nodes) {
  int nuc = 0, rev_nuc = 0, next_o = 0, length = 0, i;
  orc_node *next = NULL;
  unsigned other = node->edges;
  if ((orientation ^ 1) == 1) other >>= 4;
  if ((other & 15u) != 0) return 0; /* not a blunt end against the walk (element.c:761-774) */
  {
    int join_main_trunk = 0;
    while (orc_has_precisely_one_edge(node, orientation, &nuc)) {
      nodes[length] = node;
      next = orc_get_next_node(g, node, orientation, &next_o, nuc, &rev_nuc);
      if (!next) return -1;
      length++;
      if (length > limit) break;
      if (!orc_has_precisely_one_edge(next, next_o ^ 1, &nuc)) { join_main_trunk = 1; break; }
      node = next;
      orientation = next_o;
    }
    if (!join_main_trunk) {
      length = 0;
    } else {
      for (i = 0; i < length; i++) { nodes[i]->status = 3; nodes[i]->edges = 0; }
      orc_reset_one_edge(next, next_o ^ 1, rev_nuc);
    }
  }
  return length;
}

/* dB_graph_population.c:485-530  db_graph_db_node_prune_low_coverage (the raw
 * db_graph_get_next_node of dB_graph.c:22-62: no subgraph check) */
static int orc_prune_low_coverage(orc_graph *g, orc_node *node, uint32_t coverage) {
  int nuc, orient;
  if (node->covg > coverage) return 0;
  for (nuc = 0; nuc < 4; nuc++) {
    for (orient = 0; orient < 2; orient++) {
      if ((node->edges >> (nuc + 4 * orient)) & 1u) {
        uint64_t local[ORC_MAXN], rev[ORC_MAXN], key[ORC_MAXN];
        int i, next_o, rev_nuc;
        orc_node *next;
        for (i = 0; i < g->N; i++) local[i] = node->kmer[i];
        orc_revcomp(local, g->k, g->N, rev);
        if (orient == 1) {
          rev_nuc = (int)(local[g->N - 1] & 3);
          for (i = 0; i < g->N; i++) local[i] = rev[i];
        } else {
          rev_nuc = (int)(rev[g->N - 1] & 3);
        }
        orc_shift_insert(local, nuc, g->k, g->N);
        orc_get_key(local, g->k, g->N, key);
        next = orc_find(g, key);
        if (!next) return -1; /* the reference dereferences NULL here */
        next_o = orc_get_orientation(local, next, g->k, g->N);
        orc_reset_one_edge(next, next_o ^ 1, rev_nuc);
      }
    }
  }
  node->status = 3;
  node->edges = 0;
  return 1;
}

/* graph_operations.c:1069-1090: db_graph_clip_tips_in_union_of_all_colours
 * (dB_graph_population.c:2563-2608, limit k+1) then
 * db_graph_remove_errors_considering_covg_and_topology (:3595-3649 with
 * :3371-3462), then visited -> none (:3640).  counts[0] = tips clipped,
 * [1] = nodes pruned as tips, [2] = supernodes pruned, [3] = nodes left.
 * Returns 0, or -1 where the reference would die. */
int orc_clean(orc_graph *g, int threshold, int max_var_len, uint64_t *counts) {
  uint64_t i, cap = g->n_buckets * (uint64_t)g->W;
  int limit = 1 + g->k, rc = 0;
  orc_node **nodes = (orc_node **)calloc((size_t)limit + 2, sizeof(orc_node *));
  orc_path p;
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  for (i = 0; i < cap && rc == 0; i++) {
    orc_node *e = &g->table[i];
    int len;
    if (e->status != 1) continue;
    len = orc_clip_tip_with_orientation(g, e, 0, limit, nodes);
    if (len == 0) len = orc_clip_tip_with_orientation(g, e, 1, limit, nodes);
    if (len < 0) rc = -1;
    else if (len > 0) { counts[0]++; counts[1] += (uint64_t)len; }
  }
  free(nodes);
  p.nodes = (orc_node **)calloc((size_t)max_var_len + 2, sizeof(orc_node *));
  p.orients = (int *)calloc((size_t)max_var_len + 2, sizeof(int));
  p.labels = (int *)calloc((size_t)max_var_len + 2, sizeof(int));
  p.seq = (char *)calloc((size_t)max_var_len + 2 + (size_t)g->k, 1);
  for (i = 0; i < cap && rc == 0; i++) {
    orc_node *e = &g->table[i];
    int is_cycle = 0, length_sup, j, all_low = 1;
    if (e->status != 1) continue;
    if (!(e->covg > 0 && e->covg <= (uint32_t)threshold)) continue;
    length_sup = orc_supernode(g, e, max_var_len, &p, &is_cycle);
    if (length_sup < 0) { rc = -1; break; }
    if (length_sup <= 1 || length_sup > 2 * g->k + 2) continue;
    for (j = 1; j <= length_sup - 1 && all_low; j++)
      if (p.nodes[j]->covg > (uint32_t)threshold) all_low = 0;
    if (!all_low) continue;
    for (j = 1; j <= length_sup - 1; j++)
      if (orc_prune_low_coverage(g, p.nodes[j], (uint32_t)threshold) < 0) rc = -1;
    counts[2]++;
  }
  free(p.nodes); free(p.orients); free(p.labels); free(p.seq);
  for (i = 0; i < cap; i++) {
    if (g->table[i].status == 2) g->table[i].status = 1;
    if (g->table[i].status == 1) counts[3]++;
  }
  return rc;
}

/* print_binary_signature_NEW (file_reader.c:2930-2994) after cleaning: the
 * error-cleaning object carries tip_clipping, remv_low_cov_sups and its threshold
 * (graph_info.c:147-173, graph_operations.c:1084-1088). */
int orc_ctx_header_cleaned(int k, int N, uint32_t mean_read_len, uint64_t total_seq, int threshold,
                           uint8_t *out) {
  int n = orc_ctx_header(k, N, mean_read_len, total_seq, out);
  /* flags sit behind magic(6) ints(16) mean(4) total(8) idlen(4) id(9) seq_err(16) */
  uint8_t *flags = out + 6 + 16 + 4 + 8 + 4 + 9 + 16;
  int32_t v = threshold;
  flags[0] = 1; /* tip_clipping */
  flags[1] = 1; /* remv_low_cov_sups */
  memcpy(flags + 4, &v, 4);
  return n;
}

/* ------------------------------------------------------------------------ */
/* Branch printing, as `--mode build` runs it after the cleaning
 * (graph_operations.c:1145-1152): for every node with more than one edge on a
 * side, every neighbour on that side starts a "branch" -- the neighbour's
 * k-mer, then on through nodes with one edge each way, at most 300 of them --
 * unless an earlier branch has marked that neighbour `visited`.  Sequential,
 * order-dependent through the marks; the marks are left in the table.  The
 * product (lasv_b200/csrc/branches.cuh, branches_host.h) replays it on the
 * compressed graph; this is its checker. */

static int orc_degree(const orc_node *e, int orient) { /* element.c:1429-1447 db_node_get_degree */
  unsigned edges = e->edges;
  int j, d = 0;
  if (orient == 1) edges >>= 4;
  for (j = 0; j < 4; j++) { d += (int)(edges & 1u); edges >>= 1; }
  return d;
}

static int orc_edge_exists(const orc_node *e, int nuc, int orient) { /* element.c db_node_edge_exist */
  return (e->edges >> (nuc + 4 * orient)) & 1u;
}

/* print_branches.c:16-127  db_graph_get_sequence_for_a_single_branch_for_specific_person_or_pop
 * (seq has room for k + 400 characters; the reference's buffer is char[500]).
 * Returns the length in nodes, or -1 where the reference would die / dereference NULL. */
static int orc_single_branch(orc_graph *g, orc_node *node, int orientation, char *seq) {
  uint64_t km[ORC_MAXN];
  int length = 1, i, w, n;
  int reverse_edge = 0, next_o = 0, curr_o = 0, degree1, degree2;
  orc_node *node_ptr = NULL, *next_node = NULL;
  if (orientation == 0) for (w = 0; w < g->N; w++) km[w] = node->kmer[w];
  else orc_revcomp(node->kmer, g->k, g->N, km);
  orc_kmer_to_seq(km, g->k, g->N, seq);
  seq[g->k] = '\0';
  n = g->k;
  node->status = 2; /* visited */
  if (orc_degree(node, orientation) > 1) return length; /* :50-54 */
  if (orc_degree(node, orientation) == 0) return length; /* :56-59 */
  for (i = 0; i < 4; i++) { /* :61-73 */
    if (orc_edge_exists(node, i, orientation)) {
      next_o = 0;
      reverse_edge = 0;
      node_ptr = orc_get_next_node(g, node, orientation, &next_o, i, &reverse_edge);
      seq[n++] = "ACGT"[i];
      seq[n] = '\0';
    }
  }
  if (!node_ptr) return -1;
  curr_o = next_o;
  degree1 = orc_degree(node_ptr, 0);
  degree2 = orc_degree(node_ptr, 1);
  while (degree1 == 1 && degree2 == 1 && length <= 300) { /* :82-110 */
    int j;
    if (node_ptr->status == 2) break; /* "if in a circle" */
    length++;
    node_ptr->status = 2;
    next_node = NULL;
    for (j = 0; j < 4; j++) {
      if (orc_edge_exists(node_ptr, j, curr_o)) {
        next_o = 0;
        reverse_edge = 0;
        next_node = orc_get_next_node(g, node_ptr, curr_o, &next_o, j, &reverse_edge);
        seq[n++] = "ACGT"[j];
        seq[n] = '\0';
      }
    }
    if (!next_node) return -1; /* "Unexpected disruption ..." */
    node_ptr = next_node;
    curr_o = next_o;
    degree1 = orc_degree(node_ptr, 0);
    degree2 = orc_degree(node_ptr, 1);
  }
  if (curr_o == 0 && degree1 != 1 && degree2 == 1) { /* :112-123 */
    node_ptr->status = 2;
    length++;
  } else if (curr_o == 1 && degree2 != 1 && degree1 == 1) {
    node_ptr->status = 2;
    length++;
  } else {
    seq[--n] = '\0';
  }
  return length;
}

/* print_branches.c:131-199  db_graph_print_all_branches_for_specific_person_or_pop.
 * counts[0] = branches printed, [1] = branching nodes, [2] = nodes left `visited`.
 * keep_marks = 0 additionally resets the marks afterwards (a convenience the reference
 * does not have: there the marks stay, graph_operations.c:1145-1152).
 * Returns 0, or -1 on I/O error / inconsistent graph. */
int orc_print_branches(orc_graph *g, const char *path, int keep_marks, uint64_t *counts) {
  uint64_t i, cap = g->n_buckets * (uint64_t)g->W;
  char *seq = (char *)malloc((size_t)g->k + 512);
  int rc = 0;
  FILE *f = fopen(path, "w");
  if (!f || !seq) { free(seq); if (f) fclose(f); return -1; }
  counts[0] = counts[1] = counts[2] = 0;
  for (i = 0; i < cap; i++) /* :133 visited -> none */
    if (g->table[i].status == 2) g->table[i].status = 1;
  for (i = 0; i < cap && rc == 0; i++) {
    orc_node *cur = &g->table[i];
    int k;
    if (cur->status == 0) continue; /* an empty slot has coverage 0 */
    if (!(cur->covg > 0 && cur->status != 3 && (orc_degree(cur, 0) > 1 || orc_degree(cur, 1) > 1))) continue;
    counts[1]++;
    for (k = 0; k < 2 && rc == 0; k++) {
      int j;
      if (orc_degree(cur, k) <= 1) continue;
      for (j = 0; j < 4; j++) {
        int next_o = 0, reverse_edge = 0, length;
        orc_node *next;
        if (!orc_edge_exists(cur, j, k)) continue;
        next = orc_get_next_node(g, cur, k, &next_o, j, &reverse_edge);
        if (!next) { rc = -1; break; } /* the reference reads next_node->status of NULL */
        if (next->status == 2) continue;
        length = orc_single_branch(g, next, next_o, seq);
        if (length < 0) { rc = -1; break; }
        if (length > 0) {
          fprintf(f, ">b_%.*s\n%s\n", g->k, seq, seq);
          counts[0]++;
        }
      }
    }
  }
  for (i = 0; i < cap; i++) {
    if (g->table[i].status == 2) {
      counts[2]++;
      if (!keep_marks) g->table[i].status = 1;
    }
  }
  free(seq);
  if (fclose(f) != 0) rc = -1;
  return rc;
}
