// seq_reader.cpp -- see seq_reader.h
#include "seq_reader.h"
#include <string.h>
#include <strings.h>

namespace lasv {

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
  gz_ = gzopen(path.c_str(), "r");
  if (!gz_) {
    err = "Couldn't open sequence file '" + path + "'";
    return false;
  }
  gzbuffer(gz_, 1u << 20);
  own_.resize(4u << 20);
  buf_ = own_.data();
  memory_ = false;
  pos_ = end_ = 0;
  eof_ = false;
  n_reads_ = 0;
  // type sniffing on the first byte (seq_file.c:215-259)
  const int c = getc_();
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
    pos_--;  // unget
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

void SeqReader::close() {
  if (gz_) gzclose(gz_);
  gz_ = nullptr;
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
