// seq_reader.cpp -- see seq_reader.h
#include "seq_reader.h"
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>

namespace lasv {

// ------------------------------------------------------------------ BGZF
namespace {
struct BgzfBlock {
  size_t cdata;      // offset of the deflate stream in the file
  uint32_t clen;     // its length
  uint32_t isize;    // bytes it inflates to
  uint32_t crc;
  size_t out;        // offset in the window
};

// Header of the BGZF block at d[0..n): its total size, or 0 if this is not a BGZF block (SAM spec 4.1: gzip member
// with FLG.FEXTRA and a 'B','C' subfield of 2 bytes holding the block size minus one).
size_t bgzf_block_size(const unsigned char *d, size_t n, size_t *cdata_off) {
  if (n < 18 || d[0] != 0x1f || d[1] != 0x8b || d[2] != 8 || !(d[3] & 4)) return 0;
  const size_t xlen = (size_t)d[10] | ((size_t)d[11] << 8);
  if (12 + xlen > n) return 0;
  size_t bsize = 0;
  for (size_t p = 12; p + 4 <= 12 + xlen;) {
    const size_t slen = (size_t)d[p + 2] | ((size_t)d[p + 3] << 8);
    if (d[p] == 'B' && d[p + 1] == 'C' && slen == 2 && p + 6 <= 12 + xlen) bsize = ((size_t)d[p + 4] | ((size_t)d[p + 5] << 8)) + 1;
    p += 4 + slen;
  }
  if (!bsize || bsize > n || bsize < 12 + xlen + 8) return 0;
  // FNAME / FCOMMENT / FHCRC are not used by BGZF writers; a member that has them is left to zlib
  if (d[3] & ~4) return 0;
  *cdata_off = 12 + xlen;
  return bsize;
}
}  // namespace

BgzfStream::~BgzfStream() {
  if (data_) munmap((void *)data_, size_);
  if (fd_ >= 0) ::close(fd_);
}

bool BgzfStream::open(const char *path, int threads) {
  fd_ = ::open(path, O_RDONLY);
  if (fd_ < 0) return false;
  struct stat st;
  if (fstat(fd_, &st) != 0 || st.st_size < 28) { ::close(fd_); fd_ = -1; return false; }
  void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd_, 0);
  if (p == MAP_FAILED) { ::close(fd_); fd_ = -1; return false; }
  data_ = (const unsigned char *)p;
  size_ = (size_t)st.st_size;
  size_t off = 0;
  if (!bgzf_block_size(data_, size_, &off)) {
    munmap(p, size_);
    data_ = nullptr;
    ::close(fd_);
    fd_ = -1;
    return false;
  }
  madvise(p, size_, MADV_SEQUENTIAL);
  threads_ = threads < 1 ? 1 : threads;
  pos_ = 0;
  return true;
}

bool BgzfStream::next_impl(std::vector<unsigned char> *outv, unsigned char *dst, size_t dst_cap, size_t *n_out, std::string &err) {
  constexpr size_t WINDOW = 8u << 20;
  std::vector<BgzfBlock> blocks;
  size_t total = 0;
  while (pos_ < size_ && total < WINDOW) {
    size_t off = 0;
    const size_t bs = bgzf_block_size(data_ + pos_, size_ - pos_, &off);
    if (!bs) {
      err = "not a BGZF block where one should start (truncated or mixed gzip file)";
      return false;
    }
    const unsigned char *t = data_ + pos_ + bs - 8;
    BgzfBlock b;
    b.cdata = pos_ + off;
    b.clen = (uint32_t)(bs - off - 8);
    b.crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    b.isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
    b.out = total;
    total += b.isize;
    pos_ += bs;
    if (b.isize) blocks.push_back(b);  // (the empty block that ends a BGZF file)
  }
  if (blocks.empty()) return false;
  if (outv) {
    outv->resize(total);
    dst = outv->data();
  } else if (total > dst_cap) {
    err = "a window of BGZF blocks does not fit the caller's buffer";
    return false;
  }
  if (n_out) *n_out = total;
  std::atomic<size_t> next_block{0};
  std::atomic<int> failed{0};
  auto work = [&]() {
    MemberInflater inf;  // (this library's deflate decoder, gz_parallel.cpp: ~3x zlib's inflate on FASTQ text)
    for (;;) {
      const size_t i = next_block.fetch_add(1);
      if (i >= blocks.size() || failed) break;
      const BgzfBlock &b = blocks[i];
      if (!inf.inflate(data_ + b.cdata, b.clen, dst + b.out, b.isize) ||
          fast_crc32(0, dst + b.out, b.isize) != b.crc)
        failed = 1;
    }
  };
  const int n_thr = (int)std::min<size_t>((size_t)threads_, blocks.size());
  std::vector<std::thread> pool;
  for (int t = 1; t < n_thr; t++) pool.emplace_back(work);
  work();
  for (auto &t : pool) t.join();
  if (failed) {
    err = "corrupt BGZF block (inflate or CRC error)";
    return false;
  }
  return true;
}

static bool path_has_ext(const std::string &p, const char *ext) {
  const size_t n = strlen(ext);
  if (p.size() >= n && strcasecmp(p.c_str() + p.size() - n, ext) == 0) return true;
  // ".sam.something" style (seq_file.c:126-137 looks for the extension anywhere)
  std::string lower(p);
  for (auto &c : lower) c = (char)tolower((unsigned char)c);
  std::string needle = std::string(ext) + ".";
  return lower.find(needle) != std::string::npos;
}

bool SeqReader::open(const std::string &path, std::string &err) {
  close();
  path_ = path;
  if (path_has_ext(path, ".sam") || path_has_ext(path, ".bam")) {
    err = "SAM/BAM input is not supported by the GPU loader: " + path;
    return false;
  }
  {
    const unsigned hw = std::thread::hardware_concurrency();
    int thr = (int)std::min(16u, std::max(2u, hw / 2));  // inflate threads of a compressed file (8 on a 16-core host)
    if (const char *e = getenv("LASV_B200_GZ_THREADS")) thr = atoi(e);
    std::unique_ptr<BgzfStream> b(new BgzfStream());
    if (thr > 0 && b->open(path.c_str(), thr)) stream_ = std::move(b);
    if (thr > 0 && !stream_) {
      // ordinary gzip of some size: chunks of the compressed stream inflated by the same number of threads
      size_t min_kb = 8192, chunk_kb = 1024;
      if (const char *e = getenv("LASV_B200_PGZ_MIN_KB")) min_kb = (size_t)strtoull(e, nullptr, 10);
      if (const char *e = getenv("LASV_B200_PGZ_CHUNK_KB")) chunk_kb = (size_t)strtoull(e, nullptr, 10);
      std::unique_ptr<PgzStream> g(new PgzStream());
      if (g->open(path.c_str(), thr, min_kb << 10, chunk_kb << 10)) stream_ = std::move(g);
    }
  }
  gz_ = gzopen(path.c_str(), "r");  // (kept open in BGZF mode too: the reader's "is open" handle)
  if (!gz_) {
    err = "Couldn't open sequence file '" + path + "'";
    return false;
  }
  gzbuffer(gz_, stream_ ? 8192u : 1u << 20);
  own_.resize(stream_ ? 0 : 4u << 20);
  buf_ = own_.data();
  memory_ = false;
  pos_ = end_ = 0;
  eof_ = false;
  n_reads_ = 0;
  skip_first_ = false;
  // type sniffing on the first byte (seq_file.c:215-259).  With a stream of windows the byte comes from zlib (a few
  // KB inflated) rather than from the stream, whose first window is a round of work of all its threads; the marker is
  // then skipped when the first window arrives.
  int c;
  if (stream_) {
    unsigned char first = 0;
    c = gzread(gz_, &first, 1) == 1 ? (int)first : -1;
    skip_first_ = (c == '>' || c == '@');
  } else {
    c = getc_();
  }
  if (c == -1) {
    type_ = PLAIN;  // empty file: no reads
  } else if (c == '>') {
    type_ = FASTA;
    at_record_start_ = true;
  } else if (c == '@') {
    type_ = FASTQ;
    at_record_start_ = true;
  } else {
    type_ = PLAIN;
    if (!stream_) pos_--;  // unget
  }
  return true;
}

void SeqReader::open_memory(const unsigned char *data, size_t n, const std::string &path) {
  close();
  path_ = path;
  buf_ = data;
  memory_ = true;
  pos_ = 0;
  end_ = n;
  eof_ = false;
  n_reads_ = 0;
  type_ = FASTQ;
  at_record_start_ = false;  // the slice begins AT the '@' of a record header
}

void SeqReader::seed(std::vector<unsigned char> &&text) {
  own_ = std::move(text);
  buf_ = own_.data();
  memory_ = false;
  skip_first_ = false;
  pos_ = 0;
  end_ = own_.size();
  eof_ = false;
  at_record_start_ = false;  // the text begins AT the '@' of a record header
}

void SeqReader::close() {
  if (gz_) gzclose(gz_);
  gz_ = nullptr;
  stream_.reset();
  type_ = UNKNOWN;
  memory_ = false;
}

bool SeqReader::fill_() {
  if (eof_) return false;
  if (memory_) {  // the whole slice was the window
    eof_ = true;
    pos_ = end_ = 0;
    return false;
  }
  if (stream_) {  // the next window of text, inflated by several threads
    std::string err;
    if (!stream_->next(own_, err)) {
      if (!err.empty()) io_error_ = err + " [file: " + path_ + "]";
      eof_ = true;
      pos_ = end_ = 0;
      return false;
    }
    buf_ = own_.data();
    pos_ = skip_first_ ? 1 : 0;  // (the record marker open() has seen)
    skip_first_ = false;
    end_ = own_.size();
    return true;
  }
  const int n = gzread(gz_, own_.data(), (unsigned)own_.size());
  if (n <= 0) {
    eof_ = true;
    pos_ = end_ = 0;
    return false;
  }
  pos_ = 0;
  end_ = (size_t)n;
  return true;
}

int SeqReader::getc_() {
  if (pos_ == end_ && !fill_()) return -1;
  return buf_[pos_++];
}

int SeqReader::peekc_() {
  if (pos_ == end_ && !fill_()) return -1;
  return buf_[pos_];
}

// Consume up to and including the next '\n'; append the line, without its
// trailing "\r\n", to *dst (strbuf readline + chomp).
void SeqReader::read_line_(std::string *dst) {
  for (;;) {
    if (pos_ == end_ && !fill_()) break;
    const unsigned char *p = buf_ + pos_;
    const size_t avail = end_ - pos_;
    const unsigned char *nl = (const unsigned char *)memchr(p, '\n', avail);
    const size_t take = nl ? (size_t)(nl - p) : avail;
    if (dst) dst->append((const char *)p, take);
    pos_ += take;
    if (nl) {
      pos_++;
      break;
    }
  }
  if (dst) {
    while (!dst->empty() && (dst->back() == '\r' || dst->back() == '\n')) dst->pop_back();
  }
}

bool SeqReader::next(std::string &err) {
  const bool ok = next_record_(err);
  // the compressed stream broke off: not a clean end of file, and the cause of whatever the parser saw at the break
  if (!ok && !io_error_.empty()) err = io_error_;
  return ok;
}

bool SeqReader::next_record_(std::string &err) {
  seq_.clear();
  qual_.clear();
  if (!gz_ && !memory_) return false;
  int c;
  if (type_ == FASTQ) {
    if (!at_record_start_) {
      // skip newlines, expect '@' (seq_fastq.c:86-99)
      while ((c = getc_()) != -1 && (c == '\n' || c == '\r')) {}
      if (c == -1) return false;
      if (c != '@') {
        err = "FASTQ header does not begin with '@' [file: " + path_ + "]";
        return false;
      }
    }
    at_record_start_ = false;
    read_line_(nullptr);  // read name
    // sequence: every line up to one that starts with '+' (seq_fastq.c:24-53)
    while ((c = getc_()) != -1 && c != '+') {
      if (c != '\r' && c != '\n') {
        seq_.push_back((char)c);
        read_line_(&seq_);
      }
    }
    if (c == -1) {
      if (seq_.empty()) return false;  // trailing '@name' without a record
      err = "Missing '+' in FASTQ [file: " + path_ + "]";
      return false;
    }
    read_line_(nullptr);  // rest of the '+' line
    // qualities: exactly |seq| non-newline bytes, possibly over several lines
    qual_.reserve(seq_.size());
    while (qual_.size() < seq_.size()) {
      if (pos_ == end_ && !fill_()) {
        err = "FASTQ file ended without finishing quality scores [file: " + path_ + "]";
        return false;
      }
      // bulk: the run of quality bytes up to the next line terminator (or as many as are still wanted)
      const unsigned char *p = buf_ + pos_;
      size_t want = seq_.size() - qual_.size();
      if (want > end_ - pos_) want = end_ - pos_;
      const unsigned char *nl = (const unsigned char *)memchr(p, '\n', want);
      size_t take = nl ? (size_t)(nl - p) : want;
      const unsigned char *cr = take ? (const unsigned char *)memchr(p, '\r', take) : nullptr;
      if (cr) take = (size_t)(cr - p);
      qual_.append((const char *)p, take);
      pos_ += take;
      if (nl || cr) pos_++;  // the terminator itself is not a quality byte
    }
    n_reads_++;
    return true;
  }
  if (type_ == FASTA) {
    if (!at_record_start_) {
      // we are positioned just after a '>' or at EOF
      return false;
    }
    at_record_start_ = false;
    read_line_(nullptr);  // name
    // bases: everything but newlines up to the next '>' (seq_fasta.c:72-95)
    for (;;) {
      if (pos_ == end_ && !fill_()) break;
      const unsigned char ch = buf_[pos_];
      if (ch == '>') {
        pos_++;
        at_record_start_ = true;
        break;
      }
      if (ch == '\n' || ch == '\r') {
        pos_++;
        continue;
      }
      // take the run up to the next newline or '>'
      size_t e = pos_;
      while (e < end_ && buf_[e] != '\n' && buf_[e] != '\r' && buf_[e] != '>') e++;
      seq_.append((const char *)buf_ + pos_, e - pos_);
      pos_ = e;
    }
    n_reads_++;
    return true;
  }
  if (type_ == PLAIN) {
    // one read per line terminator (seq_plain.c:24-66): an empty line is an empty read
    if (peekc_() == -1) return false;
    while ((c = getc_()) != -1 && c != '\n' && c != '\r') seq_.push_back((char)c);
    n_reads_++;
    return true;
  }
  return false;
}

uint64_t SeqReader::fill_batch_fast(HostBatch &b, bool with_qual) {
  if (!memory_ || type_ != FASTQ || at_record_start_) return 0;
  uint64_t added = 0;
  const unsigned char *const base = buf_, *const end = buf_ + end_;
  while (pos_ < end_) {
    const unsigned char *p = base + pos_;
    if (*p != '@') break;  // blank lines before a header, or garbage: the slow path decides
    const unsigned char *h = (const unsigned char *)memchr(p, '\n', (size_t)(end - p));
    if (!h) break;
    const unsigned char *sq = h + 1;
    const unsigned char *se = sq < end ? (const unsigned char *)memchr(sq, '\n', (size_t)(end - sq)) : nullptr;
    if (!se || se == sq || se[-1] == '\r' || se + 1 >= end || se[1] != '+') break;  // empty / CRLF / multi-line sequence
    const unsigned char *pl = (const unsigned char *)memchr(se + 1, '\n', (size_t)(end - (se + 1)));
    if (!pl) break;
    const size_t len = (size_t)(se - sq);
    const unsigned char *q = pl + 1;
    if ((size_t)(end - q) < len) break;                                  // truncated quality line
    if (q + len < end && q[len] != '\n') break;                          // quality line longer / CRLF / wrapped
    if (memchr(q, '\n', len) || memchr(q, '\r', len)) break;             // wrapped quality line
    if (b.n_reads >= b.cap_reads || b.n_bytes + len + 1 > b.cap_bytes) break;  // batch full
    memcpy(b.seq + b.n_bytes, sq, len);
    b.seq[b.n_bytes + len] = '\n';
    if (with_qual) {
      memcpy(b.qual + b.n_bytes, q, len);
      b.qual[b.n_bytes + len] = '\n';
    }
    b.n_bytes += len + 1;
    b.n_reads++;
    b.offsets[b.n_reads] = b.n_bytes;
    pos_ = (size_t)(q - base) + len + (q + len < end ? 1 : 0);
    n_reads_++;
    added++;
  }
  return added;
}

bool SeqReader::append_to(HostBatch &b, bool with_qual) const {
  const uint64_t need = (uint64_t)seq_.size() + 1;
  if (b.n_reads >= b.cap_reads || b.n_bytes + need > b.cap_bytes) return false;
  memcpy(b.seq + b.n_bytes, seq_.data(), seq_.size());
  b.seq[b.n_bytes + seq_.size()] = '\n';
  if (with_qual) {
    if (qual_.size() == seq_.size()) memcpy(b.qual + b.n_bytes, qual_.data(), seq_.size());
    else memset(b.qual + b.n_bytes, 0x7F, seq_.size());  // reader without qualities in a mixed pair
    b.qual[b.n_bytes + seq_.size()] = '\n';
  }
  b.n_bytes += need;
  b.n_reads++;
  b.offsets[b.n_reads] = b.n_bytes;
  return true;
}

}  // namespace lasv
