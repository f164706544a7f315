// seq_reader.h -- host-side sequence file reader feeding the GPU loader.
//
// Re-implements the grammar of the reference's vendored reader for the formats
// laSV feeds it (libs/seq_file/seq_file.c:191-239 type sniffing on the first
// byte; seq_fastq.c:24-150; seq_fasta.c:24-100; seq_plain.c), gzip through
// zlib's gzopen like the reference (seq_file.c:170).  Reads are appended to a
// batch as two byte streams with a '\n' after each read -- the layout the build
// kernel consumes -- instead of being handed out base by base.
#pragma once
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>
#include <zlib.h>

#include "gz_parallel.h"

namespace lasv {

// Block-parallel inflate of BGZF files (the blocked gzip of bgzip / htslib that sequencing pipelines write:
// a series of gzip members of <= 64 KB, each carrying its own compressed size in a 'BC' extra field, so the members
// can be found without inflating anything).  The reference reads compressed input one byte at a time through
// gzgetc (libs/seq_file/seq_common.h:102-103, seq_file.c:170).  Ordinary single-stream gzip: PgzStream
// (gz_parallel.h); small files and anything else zlib can read: zlib's gzread, one stream per file and thread.
class BgzfStream : public TextStream {
 public:
  ~BgzfStream() override;
  // maps the file; false (and nothing kept open) unless it starts with a BGZF block
  bool open(const char *path, int threads);
  // (next / next_into: the next window of blocks, about 8 MB of text)
  void rewind() override { pos_ = 0; }
  const char *kind() const override { return "BGZF"; }

 protected:
  bool next_impl(std::vector<unsigned char> *out, unsigned char *dst, size_t dst_cap, size_t *n, std::string &err) override;

 private:
  int fd_ = -1;
  const unsigned char *data_ = nullptr;
  size_t size_ = 0, pos_ = 0;
  int threads_ = 1;
};

// Pinned (or plain) host buffers for one batch of reads.
struct HostBatch {
  uint8_t *seq = nullptr;
  uint8_t *qual = nullptr;
  uint64_t *offsets = nullptr;  // cap_reads + 1 entries
  uint64_t cap_bytes = 0, cap_reads = 0;
  uint64_t n_bytes = 0, n_reads = 0;
  // raw: `seq` holds n_bytes of plain four-line FASTQ TEXT (whole records, ending with '\n') instead of parsed
  // streams; it is taken apart on the device (fastq_parse.cu)
  bool raw = false;
  void reset() { n_bytes = 0; n_reads = 0; raw = false; if (offsets) offsets[0] = 0; }
};

class SeqReader {
 public:
  enum Type { UNKNOWN, FASTQ, FASTA, PLAIN };
  SeqReader() {}
  ~SeqReader() { close(); }
  // Returns false (and sets err) if the file cannot be opened or is SAM/BAM.
  bool open(const std::string &path, std::string &err);
  // Parse records from a byte range already in memory (a slice of an mmap'ed, uncompressed FASTQ file that
  // starts at a record header): the parallel loader gives every parser thread its own slice.
  void open_memory(const unsigned char *data, size_t n, const std::string &path);
  // True if the file is stored uncompressed (gzopen reads it transparently either way).
  bool is_plain_file() const { return !stream_ && gz_ && gzdirect(gz_) == 1; }
  // BGZF or large gzip input: the stream of windows of text inflated by several threads (the loader pulls whole
  // windows from it for the device parse), and the way back: continue parsing on the host from `text` (starts at a
  // record header) followed by the stream's next windows
  TextStream *windows() { return stream_.get(); }
  void seed(std::vector<unsigned char> &&text);
  void close();
  bool has_quality() const { return type_ == FASTQ; }
  Type type() const { return type_; }
  const std::string &path() const { return path_; }

  // Parse the next read into the internal (seq_, qual_) strings.
  // Returns false at end of file.  err is set on malformed input.
  bool next(std::string &err);
  const std::string &seq() const { return seq_; }
  const std::string &qual() const { return qual_; }
  uint64_t reads_parsed() const { return n_reads_; }

  // Fast path of the parallel loader: as many plain four-line FASTQ records ('@name', one sequence line, '+...',
  // one quality line of the same length, LF line ends) as fit, straight from an in-memory slice into the batch
  // -- one copy, no intermediate strings.  Stops in front of the first record that does not have that shape
  // (next() handles it, with the reference's full grammar) or does not fit.  Returns the number of reads added.
  uint64_t fill_batch_fast(HostBatch &b, bool with_qual);

  // Append the current read to a batch; false if it does not fit.
  bool append_to(HostBatch &b, bool with_qual) const;

 private:
  bool next_record_(std::string &err);
  int getc_();
  int peekc_();
  bool fill_();
  void read_line_(std::string *dst);  // rest of the current line, chomped; dst may be null

  gzFile gz_ = nullptr;
  std::unique_ptr<TextStream> stream_;  // BGZF / large gzip input: windows of text inflated by several threads
  std::string path_;
  Type type_ = UNKNOWN;
  std::vector<unsigned char> own_;      // gz mode: the refill buffer
  const unsigned char *buf_ = nullptr;  // current window (own_.data() or the caller's memory)
  bool memory_ = false;
  size_t pos_ = 0, end_ = 0;
  bool eof_ = false;
  std::string io_error_;                // a read error of the compressed stream (reported by next())
  bool at_record_start_ = false;  // header marker already consumed (seq_file.c: read_line_start)
  bool skip_first_ = false;       // stream mode: ... by open(), from zlib; the first window still holds it
  std::string seq_, qual_;
  uint64_t n_reads_ = 0;
};

}  // namespace lasv
