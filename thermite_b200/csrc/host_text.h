// host_text.h -- append-only text buffers of the PAF / SAM writers (host_io.cpp, host_stream.cpp) and the internal
// interfaces between them.  Product code: must never include anything from oracle/.
#pragma once
#include <cstdlib>
#include <cstring>
#include <string>

#include "tg_internal.h"

// Append-only text buffer (std::string's per-character push_back dominated the writers' time).
struct TgText {
  char* p = nullptr;
  size_t n = 0, cap = 0;
  TgText() = default;
  TgText(const TgText&) = delete;
  TgText& operator=(const TgText&) = delete;
  TgText(TgText&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
  ~TgText() { free(p); }
  void reserve(size_t c) {
    if (c <= cap) return;
    size_t want = cap ? cap : 4096;
    while (want < c) want += want / 2 + 4096;
    p = (char*)realloc(p, want);
    cap = want;
  }
  char* room(size_t k) { if (n + k > cap) reserve(n + k); return p + n; }
  void append(const char* src, size_t k) { memcpy(room(k), src, k); n += k; }
  void push_back(char c) { *room(1) = c; n++; }
  TgText& operator+=(const char* lit) { append(lit, strlen(lit)); return *this; }
  TgText& operator+=(const std::string& str) { append(str.data(), str.size()); return *this; }
  size_t size() const { return n; }
  const char* data() const { return p; }
};
struct TgOut {
  TgText s;
  void num(uint64_t v) {  // decimal digits, two at a time
    static const char lut[201] =
        "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263"
        "646566676869707172737475767778798081828384858687888990919293949596979899";
    char buf[24];
    int k = 24;
    while (v >= 100) {
      const unsigned q = (unsigned)(v % 100);
      v /= 100;
      buf[--k] = lut[2 * q + 1]; buf[--k] = lut[2 * q];
    }
    if (v >= 10) { buf[--k] = lut[2 * v + 1]; buf[--k] = lut[2 * v]; }
    else buf[--k] = (char)('0' + v);
    memcpy(s.room((size_t)(24 - k)), buf + k, (size_t)(24 - k));
    s.n += (size_t)(24 - k);
  }
  void snum(int64_t v) {
    if (v < 0) { s.push_back('-'); num((uint64_t)(-v)); } else num((uint64_t)v);
  }
};


// The records of a batch as the writers see them: wide (tg_result) or compact (tg_result_c) records behind one view.
struct TgRecView {
  const uint64_t* first64 = nullptr;  // wide results
  const uint32_t* first32 = nullptr;  // compact results
  const uint32_t* count = nullptr;
  const tg_aln* wide = nullptr;
  const tg_aln_c* comp = nullptr;
  const uint32_t* ops = nullptr;
};
// records of the reads [r0, r1) as PAF (sam = 0) or SAM (sam = 1) text, appended to o (host_io.cpp)
void tg_format_reads(const tg_index_host* ix, const TgRecView& v, const uint8_t* bases, const uint64_t* offs, const uint8_t* names,
                     const uint64_t* name_offs, const uint8_t* quals, const uint64_t* qual_offs, int sam, uint32_t r0, uint32_t r1,
                     TgOut& o);
// FASTQ text, two passes so that many threads can write one batch in place (host_io.cpp).  A segment starts at a record start.
struct TgFastqCount {
  uint64_t n = 0, bases = 0, names = 0, quals = 0;
  size_t consumed = 0;  // offset (in the whole text) just behind the last complete record of the segment
  bool bad = false;
};
void tg_fastq_count(const char* text, size_t begin, size_t end, bool final, TgFastqCount& c);
void tg_fastq_fill(const char* text, size_t begin, size_t end, uint64_t n, uint8_t* bases, uint64_t* offs, uint64_t base0,
                   uint8_t* names, uint64_t* name_offs, uint64_t name0, uint8_t* quals, uint64_t* qual_offs, uint64_t qual0);
size_t tg_fastq_record_start(const char* text, size_t len, size_t p);
// offset just behind the k-th complete record of [begin, end) (end when there are fewer)
size_t tg_fastq_skip(const char* text, size_t begin, size_t end, uint64_t k);
// SAM lines -> BGZF-compressed BAM records (host_bam.cpp)
tg_status tg_sam_text_to_bam(const tg_index_host* ix, const char* sam, size_t sl, bool append_eof, std::string& z, unsigned max_threads = 0);
