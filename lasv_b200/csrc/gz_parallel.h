// gz_parallel.h -- windows of inflated text from compressed sequence files, inflated by several threads.
//
// The reference reads compressed input through zlib's gzopen, one byte at a time (libs/seq_file/seq_file.c:170,
// seq_common.h:102-103: gzgetc).  Two kinds of gzip file can be inflated by several threads at once:
//   BgzfStream  blocked gzip (bgzip / htslib): gzip members of <= 64 KB that carry their compressed size in a 'BC'
//               extra field, so the members are found without inflating anything (seq_reader.cpp)
//   PgzStream   ordinary single-stream gzip (gzip, pigz, what most sequencers' software writes).  The deflate stream
//               has no index, so the file is cut into chunks of compressed bytes and every thread but the first
//               starts in the middle of the stream: it searches its chunk for the start of a deflate block (a
//               dynamic-Huffman header whose code lengths form complete codes and whose block inflates to text),
//               and inflates from there WITHOUT knowing the 32 KB of text in front of it -- into 16-bit symbols,
//               where a value >= 256 stands for "byte number v-256 of the unknown window" and back-references copy
//               symbols like bytes.  The thread of the chunk before it inflates until it stands at a block boundary
//               at or behind that start.  If the two meet on the same bit the guess was a real boundary (the
//               decoder is deterministic given the position and the window), the window becomes known from the
//               earlier chunk's last 32 KB and the symbols are translated to bytes; if the earlier thread steps
//               over the guess it was not a boundary, the work of that chunk is dropped and the earlier thread
//               carries on through it.  Either way the text is, byte for byte, what a sequential inflate gives,
//               and CRC-32 and length of every gzip member are checked as zlib does.
//               (The two-pass scheme follows Kerbiriou & Chikhi, "Parallel decompression of gzip-compressed files
//               and random access to DNA sequences", 2019; the decoder here is written for this library.)
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>

namespace lasv {

// gzip's CRC-32, continued from `crc` like zlib's crc32() (carry-less multiplication where the CPU has it)
uint32_t fast_crc32(uint32_t crc, const unsigned char *buf, size_t len);

class TextStream {
 public:
  virtual ~TextStream() {}
  // the next window of text (at most 8 MB + 64 KB) into `out`; false at the end of the file or on error (then err
  // is set)
  bool next(std::vector<unsigned char> &out, std::string &err) { return next_impl(&out, nullptr, 0, nullptr, err); }
  // the same straight into the caller's buffer (dst_cap >= 8 MB + 64 KB): *n receives the bytes written
  bool next_into(unsigned char *dst, size_t dst_cap, size_t *n, std::string &err) { return next_impl(nullptr, dst, dst_cap, n, err); }
  // back to the start of the file
  virtual void rewind() = 0;
  virtual const char *kind() const = 0;

 protected:
  virtual bool next_impl(std::vector<unsigned char> *out, unsigned char *dst, size_t dst_cap, size_t *n, std::string &err) = 0;
};

// A raw deflate stream whose text depends on nothing in front of it (the body of a gzip member, e.g. a BGZF block),
// inflated to bytes by this library's decoder.  One per thread.
class MemberInflater {
 public:
  MemberInflater();
  ~MemberInflater();
  // src[0, n) must inflate to exactly `expect` bytes, which are written to dst; false if it does not (corrupt data)
  bool inflate(const unsigned char *src, size_t n, unsigned char *dst, size_t expect);

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

class PgzStream : public TextStream {
 public:
  PgzStream();
  ~PgzStream() override;
  // maps the file; false (and nothing kept open) unless it is a gzip file (deflate) of at least min_bytes
  bool open(const char *path, int threads, size_t min_bytes, size_t chunk_bytes);
  void rewind() override;
  const char *kind() const override { return "gzip"; }
  // chunks whose guessed start was confirmed / dropped so far (the tests and the speed tool look at these)
  uint64_t chunks_confirmed() const { return confirmed_; }
  uint64_t chunks_dropped() const { return dropped_; }

 protected:
  bool next_impl(std::vector<unsigned char> *out, unsigned char *dst, size_t dst_cap, size_t *n, std::string &err) override;

 private:
  struct Chunk;
  bool run_round(std::string &err);
  void reset();
  void recycle(std::unique_ptr<Chunk> &c);

  int fd_ = -1;
  const unsigned char *data_ = nullptr;
  size_t size_ = 0;
  int threads_ = 1;
  size_t chunk_bytes_ = 1u << 20;
  // where the next round starts: a block boundary (or the start of a member's deflate stream) and the text before it
  uint64_t cur_bit_ = 0;
  std::vector<unsigned char> window_;  // 32 KB; the last window_valid_ bytes are text of the current member
  size_t window_valid_ = 0;
  bool finished_ = false;              // the stream has ended (or broke: pending_err_)
  bool sequential_ = false;            // two rounds without a confirmed chunk start: one chunk per round from now on
  int failed_rounds_ = 0;
  std::string pending_err_;
  // the confirmed chunks of the last round, in stream order, and how far they have been handed out
  std::vector<std::unique_ptr<Chunk>> queue_;
  size_t q_chunk_ = 0, q_off_ = 0, q_end_ = 0;  // (q_end_: index of the next member end of queue_[q_chunk_])
  uint32_t crc_run_ = 0, len_run_ = 0;          // CRC-32 / length of the current member so far
  uint64_t confirmed_ = 0, dropped_ = 0;
  std::vector<std::pair<uint16_t *, size_t>> pool_;  // symbol buffers of finished chunks, for the next round
};

}  // namespace lasv
