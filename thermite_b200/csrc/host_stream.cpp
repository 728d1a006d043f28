// host_stream.cpp -- streaming FASTQ ingest (plain, gzip, BGZF) and the file-to-file pipeline of libthermite_gpu.
//
// Replaces the read loop of align_reads_from_file (reference src/aligner.rs:22-120): needletail::parse_fastx_file is
// gz-transparent and yields one record at a time (src/aligner.rs:51-55); the reference's own data sets are .fastq.gz
// (data/Makefile:26,35).  Here a reader hands out BATCHES of reads in page-locked buffers -- text is inflated (BGZF blocks
// on all host cores, plain gzip members on one), cut at record starts and parsed by all host cores straight into the
// batch -- and tg_align_files runs reader, aligner and writers as three overlapped stages, so that no stage ever holds a
// whole file and the GPU works on batch k while batch k + 1 is parsed and batch k - 1 is written.
// Product code: must never include anything from oracle/.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include "host_text.h"
#include "tg_internal.h"

namespace {

// threads a stage may use: all cores by default; TG_IO_THREADS overrides (reader and writers run at the same time inside
// tg_align_files, so half the cores each avoids oversubscription there)
unsigned host_threads() {
  static const unsigned n = []() {
    const char* e = getenv("TG_IO_THREADS");
    const long v = e ? atol(e) : 0;
    return v > 0 ? (unsigned)std::min<long>(v, 64) : (unsigned)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 64);
  }();
  return n;
}

// Parser / formatter workers run at a lower priority than the thread that drives the GPU: with reader and writers busy
// on every core, the launching thread was descheduled between kernel launches and a 25 ms batch took 50 ms.
void lower_priority() { setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 10); }

template <class F>
void run_threads(unsigned T, F&& f) {
  if (T <= 1) { f(0u); return; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; t++) th.emplace_back([&f, t]() { lower_priority(); f(t); });
  for (auto& x : th) x.join();
}

// Grow-only buffer, page-locked when a CUDA device is there (the batch goes to the GPU by DMA), pageable otherwise (the
// reader is host logic and is tested without a GPU).
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  ~HostBuf() { release(); }
  void release() {
    if (!p) return;
    if (pinned) cudaFreeHost(p); else free(p);
    p = nullptr; cap = 0;
  }
  bool ensure(size_t bytes, bool want_pinned) {
    if (bytes <= cap) return true;
    release();
    size_t want = bytes + bytes / 8 + 4096;
    if (want_pinned && cudaHostAlloc(&p, want, cudaHostAllocPortable) == cudaSuccess) { pinned = true; cap = want; return true; }
    if (want_pinned) cudaGetLastError();  // no device / no driver: clear the sticky error and fall back to pageable memory
    p = malloc(want);
    pinned = false;
    cap = p ? want : 0;
    return p != nullptr;
  }
};

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ---- one segment of FASTQ text: a line index built in ONE pass, records counted and copied from it ---------------------------
// (The first version walked every segment twice with memchr, once to count and once to copy: 8 calls on lines of 10-90
// bytes per record.  Here the line terminators are found 32 bytes per step and both passes work from their offsets.)
// Record rules as tg_parse_fastq: four lines per record, blank lines between records skipped, '\r' dropped, a record is
// complete when its quality line is terminated or the text ends for good.
struct SegIndex {
  size_t begin = 0, end = 0;
  std::vector<uint32_t> nl;   // offset (from begin) of every line terminator; an unterminated last line of a final text gets end
  std::vector<uint32_t> rec;  // index of the header line of every complete record
  uint64_t n = 0, bases = 0, names = 0, quals = 0;
  size_t consumed = 0;        // offset in the whole text just behind the last complete record (and blank lines before it)
  bool bad = false, has_cr = false;
  void scan(const char* text, bool final) {
    nl.clear();
    const size_t len = end - begin;
    nl.reserve(len / 40 + 16);
    const char* p = text + begin;
    size_t i = 0;
    has_cr = false;
#if defined(__AVX2__)
    const __m256i nlv = _mm256_set1_epi8('\n'), crv = _mm256_set1_epi8('\r');
    __m256i any_cr = _mm256_setzero_si256();
    for (; i + 32 <= len; i += 32) {
      const __m256i v = _mm256_loadu_si256((const __m256i*)(p + i));
      uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, nlv));
      any_cr = _mm256_or_si256(any_cr, _mm256_cmpeq_epi8(v, crv));
      while (m) { nl.push_back((uint32_t)(i + (size_t)__builtin_ctz(m))); m &= m - 1; }
    }
    has_cr = _mm256_movemask_epi8(any_cr) != 0;
#endif
    for (; i < len; i++) { if (p[i] == '\n') nl.push_back((uint32_t)i); else if (p[i] == '\r') has_cr = true; }
    if (final && len > 0 && p[len - 1] != '\n') nl.push_back((uint32_t)len);
  }
  // line k: [lb, le) relative to begin, '\r' stripped
  inline void line(const char* text, size_t k, uint32_t& lb, uint32_t& le) const {
    lb = k == 0 ? 0u : nl[k - 1] + 1u;
    le = nl[k];
    if (has_cr && le > lb && text[begin + le - 1] == '\r') le--;  // (no '\r' anywhere in the segment: the text is not touched)
  }
  size_t after_line(size_t k) const { const size_t o = begin + (size_t)nl[k] + 1; return o < end ? o : end; }
  // records among the first max_records; sums and `consumed` for exactly those
  void count(const char* text, uint64_t max_records) {
    rec.clear();
    n = bases = names = quals = 0; bad = false; consumed = begin;
    const size_t L = nl.size();
    size_t k = 0;
    while (k < L && n < max_records) {
      uint32_t lb, le;
      line(text, k, lb, le);
      if (lb == le) { consumed = after_line(k); k++; continue; }  // blank line between records
      if (k + 3 >= L) break;                                        // incomplete last record
      if (text[begin + lb] != '@') { bad = true; return; }
      uint32_t sb, se, qb, qe;
      line(text, k + 1, sb, se);
      line(text, k + 3, qb, qe);
      rec.push_back((uint32_t)k);
      n++; names += le - lb - 1; bases += se - sb; quals += qe - qb;
      consumed = after_line(k + 3);
      k += 4;
    }
  }
  void fill(const char* text, uint8_t* bases_o, uint64_t* offs, uint64_t b, uint8_t* names_o, uint64_t* name_offs, uint64_t nm,
            uint8_t* quals_o, uint64_t* qual_offs, uint64_t q) const {
    for (size_t i = 0; i < rec.size(); i++) {
      const size_t k = rec[i];
      uint32_t lb, le, sb, se, qb, qe;
      line(text, k, lb, le); line(text, k + 1, sb, se); line(text, k + 3, qb, qe);
      memcpy(names_o + nm, text + begin + lb + 1, le - lb - 1); nm += le - lb - 1;
      memcpy(bases_o + b, text + begin + sb, se - sb); b += se - sb;
      memcpy(quals_o + q, text + begin + qb, qe - qb); q += qe - qb;
      offs[i + 1] = b; name_offs[i + 1] = nm; qual_offs[i + 1] = q;  // offs / name_offs / qual_offs point at the segment's first record
    }
  }
};

// ---- reader ---------------------------------------------------------------------------------------------------------
struct tg_fastq_reader {
  FILE* f = nullptr;
  int format = 0;  // 0 plain, 1 gzip (serial inflate), 2 BGZF (blocks inflated in parallel)
  std::vector<unsigned char> zin;  // compressed bytes not yet inflated
  size_t zin_len = 0;
  bool file_eof = false;
  z_stream zs;
  bool zs_open = false;
  std::vector<char> text;  // inflated / plain text not yet handed out; starts at a record start
  size_t text_len = 0;
  int pread_fd = -1;          // a plain regular file: read with pread from all cores
  size_t file_size = 0, file_off = 0;
  const char* tptr() const { return text.data(); }
  bool text_final = false;  // nothing will be appended to `text` any more
  double bytes_per_read = 220.0;
  bool pin_quals = false;     // qualities in page-locked memory too (they go to the GPU when it writes SAM: tg_align_files)
  struct Set {
    HostBuf bases, offs, names, name_offs, quals, qual_offs;
  } set[3];
  int cur = 0;
  uint64_t total_reads = 0, total_text = 0;
  std::vector<SegIndex> seg;  // per-thread segments of the current batch (their vectors are reused from batch to batch)
  std::string err;

  ~tg_fastq_reader() {
    if (zs_open) inflateEnd(&zs);
    if (f) fclose(f);
  }
  void grow_text(size_t need) {
    if (text.size() < need) text.resize(need + need / 4 + (1u << 20));
  }
  bool read_file(size_t want) {  // appends up to `want` bytes of the file to zin; false on a read error
    if (file_eof) return true;
    if (zin.size() < zin_len + want) zin.resize(zin_len + want);
    const size_t got = fread(zin.data() + zin_len, 1, want, f);
    if (got < want) {
      if (ferror(f)) { err = "read error"; return false; }
      file_eof = true;
    }
    zin_len += got;
    return true;
  }
  bool fill_plain(size_t target) {
    // Regular file: every core reads its slice of the next stretch straight into the text buffer.  (Mapping the file and
    // parsing in place was tried first: no copy, but 48 k first-touch page faults per 1 M reads, taken by sixteen parser
    // threads under one lock, cost more than the copy; one thread's fread cost 40 ms per 1 M reads.)
    while (pread_fd >= 0 && text_len < target && !text_final) {
      size_t want = std::min(std::max<size_t>(target - text_len, 1u << 20), file_size - file_off);
      if (want == 0) { text_final = true; break; }
      grow_text(text_len + want);
      char* dst = text.data() + text_len;
      const unsigned T = want < (4u << 20) ? 1u : host_threads();
      std::vector<int> bad(T, 0);
      run_threads(T, [&](unsigned t) {
        size_t a = want * t / T;
        const size_t b = want * (t + 1) / T;
        while (a < b) {
          const ssize_t g = pread(pread_fd, dst + a, b - a, (off_t)(file_off + a));
          if (g <= 0) { bad[t] = 1; return; }  // (the file shrank under us, or an I/O error)
          a += (size_t)g;
        }
      });
      for (int x : bad) if (x) { err = "read error"; return false; }
      text_len += want; file_off += want;
      if (file_off >= file_size) text_final = true;
    }
    while (text_len < target && !text_final) {
      const size_t want = std::max<size_t>(target - text_len, 1u << 20);
      grow_text(text_len + want);
      const size_t got = fread(text.data() + text_len, 1, want, f);
      if (got < want) {
        if (ferror(f)) { err = "read error"; return false; }
        text_final = true;
      }
      text_len += got;
    }
    return true;
  }
  // gzip, any number of members (RFC 1952: a file is a series of members; needletail / flate2's MultiGzDecoder read them all)
  bool fill_gzip(size_t target) {
    size_t zpos = 0;
    while (text_len < target && !text_final) {
      if (zpos == zin_len) {
        zin_len = 0; zpos = 0;
        if (!read_file(4u << 20)) return false;
        if (zin_len == 0) {  // end of the file
          if (zs_open) { err = "gzip stream ends inside a member"; return false; }
          text_final = true;
          break;
        }
      }
      if (!zs_open) {
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, 15 + 32) != Z_OK) { err = "inflateInit2 failed"; return false; }
        zs_open = true;
      }
      grow_text(text_len + (8u << 20));
      zs.next_in = zin.data() + zpos;
      zs.avail_in = (uInt)std::min<size_t>(zin_len - zpos, 1u << 30);
      zs.next_out = (Bytef*)text.data() + text_len;
      zs.avail_out = (uInt)std::min<size_t>(text.size() - text_len, 1u << 30);
      const uInt in0 = zs.avail_in, out0 = zs.avail_out;
      const int rc = inflate(&zs, Z_NO_FLUSH);
      zpos += in0 - zs.avail_in;
      text_len += out0 - zs.avail_out;
      if (rc == Z_STREAM_END) { inflateEnd(&zs); zs_open = false; }
      else if (rc != Z_OK && rc != Z_BUF_ERROR) { err = std::string("gzip data error: ") + (zs.msg ? zs.msg : "?"); return false; }
    }
    // keep what was not consumed
    if (zpos < zin_len) memmove(zin.data(), zin.data() + zpos, zin_len - zpos);
    zin_len -= zpos;
    return true;
  }
  // BGZF (SAM spec 4.1): gzip members of at most 64 KiB, each with a 'BC' extra subfield holding its size, so the
  // members of a chunk are independent jobs
  struct Block { size_t src, clen, dst; uint32_t isize; };
  bool fill_bgzf(size_t target) {
    while (text_len < target && !text_final) {
      if (!read_file(std::max<size_t>((target - text_len) / 3, 4u << 20))) return false;
      if (zin_len == 0) { text_final = true; break; }
      std::vector<Block> blocks;
      size_t p = 0, out_bytes = 0;
      bool not_bgzf = false;
      while (p + 18 <= zin_len) {
        const unsigned char* h = zin.data() + p;
        if (!(h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4))) { not_bgzf = true; break; }
        const size_t xlen = h[10] | (h[11] << 8);
        if (p + 12 + xlen > zin_len) break;
        size_t bsize = 0;
        for (size_t q = 12; q + 4 <= 12 + xlen;) {
          const size_t slen = h[q + 2] | (h[q + 3] << 8);
          if (h[q] == 'B' && h[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) bsize = (size_t)(h[q + 4] | (h[q + 5] << 8)) + 1;
          q += 4 + slen;
        }
        if (bsize == 0) { not_bgzf = true; break; }
        if (bsize < 12 + xlen + 8) { err = "corrupt BGZF block"; return false; }
        if (p + bsize > zin_len) break;  // incomplete block: wait for more input
        const unsigned char* tail = h + bsize - 4;
        const uint32_t isize = tail[0] | (tail[1] << 8) | (tail[2] << 16) | ((uint32_t)tail[3] << 24);
        blocks.push_back(Block{p + 12 + xlen, bsize - (12 + xlen) - 8, out_bytes, isize});
        out_bytes += isize;
        p += bsize;
      }
      if (blocks.empty() && !not_bgzf) {
        if (file_eof) { if (zin_len) { err = "BGZF stream ends inside a block"; return false; } text_final = true; }
        continue;
      }
      if (!blocks.empty()) {
        grow_text(text_len + out_bytes);
        char* dst0 = text.data() + text_len;
        const unsigned T = (unsigned)std::min<size_t>(host_threads(), blocks.size());
        std::vector<int> bad(T, 0);
        run_threads(T, [&](unsigned t) {
          z_stream z;
          memset(&z, 0, sizeof(z));
          if (inflateInit2(&z, -15) != Z_OK) { bad[t] = 1; return; }
          const size_t b0 = blocks.size() * t / T, b1 = blocks.size() * (t + 1) / T;
          for (size_t b = b0; b < b1; b++) {
            const Block& k = blocks[b];
            z.next_in = zin.data() + k.src; z.avail_in = (uInt)k.clen;
            z.next_out = (Bytef*)dst0 + k.dst; z.avail_out = k.isize;
            const int rc = inflate(&z, Z_FINISH);
            if (rc != Z_STREAM_END || z.avail_out != 0) { bad[t] = 1; break; }
            inflateReset(&z);
          }
          inflateEnd(&z);
        });
        for (int b : bad) if (b) { err = "corrupt BGZF block"; return false; }
        text_len += out_bytes;
      }
      if (p < zin_len) memmove(zin.data(), zin.data() + p, zin_len - p);
      zin_len -= p;
      if (not_bgzf) { format = 1; return fill_gzip(target); }  // a plain gzip member in between: go on serially
    }
    return true;
  }
  bool fill_text(size_t target) {
    if (format == 0) return fill_plain(target);
    if (format == 1) return fill_gzip(target);
    return fill_bgzf(target);
  }
};

extern "C" {

tg_status tg_fastq_open(const char* path, tg_fastq_reader** out) {
  TG_GUARD_BEGIN
  if (!path || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  FILE* f = strcmp(path, "-") == 0 ? stdin : fopen(path, "rb");
  if (!f) return tg_fail(TG_ERR_IO, std::string("cannot open ") + path);
  auto* r = new tg_fastq_reader();
  r->f = f;
  // format from the first bytes (needletail sniffs the magic as well): 1f 8b = gzip; FEXTRA + a 'BC' subfield = BGZF
  if (!r->read_file(64u << 10)) { delete r; return tg_fail(TG_ERR_IO, "read error"); }
  if (r->zin_len >= 2 && r->zin[0] == 0x1f && r->zin[1] == 0x8b) {
    r->format = 1;
    if (r->zin_len >= 18 && r->zin[2] == 8 && (r->zin[3] & 4)) {
      const size_t xlen = r->zin[10] | (r->zin[11] << 8);
      for (size_t q = 12; q + 6 <= 12 + xlen && q + 6 <= r->zin_len;) {
        const size_t slen = r->zin[q + 2] | (r->zin[q + 3] << 8);
        if (r->zin[q] == 'B' && r->zin[q + 1] == 'C' && slen == 2) { r->format = 2; break; }
        q += 4 + slen;
      }
    }
  } else {  // plain text from a pipe: what was read is text already
    r->grow_text(r->zin_len);
    memcpy(r->text.data(), r->zin.data(), r->zin_len);
    r->text_len = r->zin_len;
    r->zin_len = 0;
    if (r->file_eof) r->text_final = true;
    // a regular file is read on from here with positional reads from all cores (fill_plain); a pipe through the stream
    struct stat sb;
    if (f != stdin && fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode)) {
      r->pread_fd = fileno(f); r->file_size = (size_t)sb.st_size; r->file_off = r->text_len;
    }
  }
  *out = r;
  return TG_OK;
  TG_GUARD_END
}

int tg_fastq_format(const tg_fastq_reader* r) { return r ? r->format : -1; }

void tg_fastq_close(tg_fastq_reader* r) {
  if (!r) return;
  if (r->f == stdin) r->f = nullptr;
  delete r;
}

tg_status tg_fastq_next(tg_fastq_reader* r, uint32_t max_reads, tg_read_batch* out) {
  TG_GUARD_BEGIN
  if (!r || !out || max_reads == 0) return tg_fail(TG_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  const unsigned T0 = host_threads();
  size_t target = (size_t)((double)max_reads * r->bytes_per_read * 1.03) + (64u << 10);
  std::vector<SegIndex>& seg = r->seg;
  uint64_t n = 0;
  size_t consumed = 0;
  size_t used = 0;
  static const bool timing = getenv("TG_READER_TIMING") != nullptr;
  double tq0 = now_ms(), tq1 = 0, tq2 = 0;
  for (;;) {
    if (!r->fill_text(target)) return tg_fail(TG_ERR_IO, "FASTQ input: " + r->err);
    tq1 = now_ms();
    const char* text = r->tptr();
    const size_t len = std::min(r->text_len, target);  // (one batch's worth of the text that is there)
    const bool at_end = r->text_final && len == r->text_len;
    // segments cut at record starts; every core indexes the lines of one and counts its records
    unsigned T = len < (4u << 20) ? 1u : T0;
    std::vector<size_t> cut(T + 1, len);
    cut[0] = 0;
    for (unsigned t = 1; t < T; t++) cut[t] = std::max(cut[t - 1], tg_fastq_record_start(text, len, (size_t)((double)len * t / T)));
    if (seg.size() < T) seg.resize(T);
    run_threads(T, [&](unsigned t) {
      seg[t].begin = cut[t]; seg[t].end = cut[t + 1];
      // an inner segment ends at a record start, so its last line is terminated; only the last segment can be cut short
      seg[t].scan(text, t + 1 < T || at_end);
      seg[t].count(text, ~0ull);
    });
    n = 0; consumed = 0; used = 0;
    for (unsigned t = 0; t < T; t++) {
      if (seg[t].bad) return tg_fail(TG_ERR_IO, "FASTQ record does not start with '@'");
      if (seg[t].begin == seg[t].end) { used = t + 1; continue; }
      if (n + seg[t].n > max_reads) {  // the batch ends inside this segment: count again, up to the cut
        seg[t].count(text, max_reads - n);
        n += seg[t].n;
        consumed = seg[t].consumed;
        used = t + 1;
        break;
      }
      n += seg[t].n;
      consumed = seg[t].consumed;
      used = t + 1;
      if (seg[t].consumed < seg[t].end && t + 1 < T) {
        // cannot happen for an inner segment (it ends at a record start); guard against a cut heuristic gone wrong
        return tg_fail(TG_ERR_IO, "FASTQ text: records are not four lines each");
      }
    }
    if (n > 0 || at_end) break;
    target = std::max(target * 2, len + (1u << 20));  // not a single complete record yet: read on
  }
  if (n == 0) {  // end of input (a truncated last record is dropped, as needletail's reader does on EOF)
    r->text_len = 0;
    return TG_OK;
  }
  tq2 = now_ms();
  uint64_t nb = 0, nn = 0, nq = 0;
  for (size_t t = 0; t < used; t++) { nb += seg[t].bases; nn += seg[t].names; nq += seg[t].quals; }
  tg_fastq_reader::Set& S = r->set[r->cur];
  r->cur = (r->cur + 1) % 3;
  if (!S.bases.ensure(nb + 64, true) || !S.offs.ensure((n + 1) * 8, true) || !S.names.ensure(nn + 1, true) ||
      !S.name_offs.ensure((n + 1) * 8, true) || !S.quals.ensure(nq + 1, r->pin_quals) || !S.qual_offs.ensure((n + 1) * 8, r->pin_quals))
    return tg_fail(TG_ERR_INTERNAL, "out of memory");
  uint8_t* bases = (uint8_t*)S.bases.p; uint64_t* offs = (uint64_t*)S.offs.p;
  uint8_t* names = (uint8_t*)S.names.p; uint64_t* name_offs = (uint64_t*)S.name_offs.p;
  uint8_t* quals = (uint8_t*)S.quals.p; uint64_t* qual_offs = (uint64_t*)S.qual_offs.p;
  offs[0] = 0; name_offs[0] = 0; qual_offs[0] = 0;
  std::vector<uint64_t> r0(used), b0(used), n0(used), q0(used);
  {
    uint64_t cr = 0, cb = 0, cn = 0, cq = 0;
    for (size_t t = 0; t < used; t++) {
      r0[t] = cr; b0[t] = cb; n0[t] = cn; q0[t] = cq;
      cr += seg[t].n; cb += seg[t].bases; cn += seg[t].names; cq += seg[t].quals;
    }
  }
  const char* text = r->tptr();
  run_threads((unsigned)used, [&](unsigned t) {
    if (seg[t].n == 0) return;
    seg[t].fill(text, bases, offs + r0[t], b0[t], names, name_offs + r0[t], n0[t], quals, qual_offs + r0[t], q0[t]);
  });
  const double tq3 = now_ms();
  // what is left starts at a record start
  if (consumed < r->text_len) memmove(r->text.data(), r->text.data() + consumed, r->text_len - consumed);
  r->text_len -= consumed;
  r->bytes_per_read = 0.5 * r->bytes_per_read + 0.5 * ((double)consumed / (double)n);
  r->total_reads += n; r->total_text += consumed;
  if (timing) fprintf(stderr, "[reader] %llu reads: fill %.1f ms, index %.1f ms, copy out %.1f ms, keep rest (%zu B) %.1f ms\n", (unsigned long long)n, tq1 - tq0, tq2 - tq1, tq3 - tq2, r->text_len, now_ms() - tq3);
  out->n_reads = (uint32_t)n;
  out->bases = bases; out->offs = offs; out->names = names; out->name_offs = name_offs; out->quals = quals; out->qual_offs = qual_offs;
  return TG_OK;
  TG_GUARD_END
}

}  // extern "C"

// ---- file-to-file pipeline ---------------------------------------------------------------------------------------------
namespace {

struct Job {
  tg_read_batch batch;
  tg_result_c res;
  const char* text = nullptr;  // PAF written on the GPU (tg_paf_*): the batch's text instead of its records
  size_t text_len = 0;
  uint64_t index = 0;
  bool end = false;
};
// bounded hand-over between two stages
struct Channel {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job> q;
  void push(const Job& j) {
    { std::lock_guard<std::mutex> l(mu); q.push_back(j); }
    cv.notify_all();
  }
  Job pop() {
    std::unique_lock<std::mutex> l(mu);
    cv.wait(l, [&] { return !q.empty(); });
    Job j = q.front();
    q.pop_front();
    return j;
  }
};
struct Progress {  // batches completely written; the earlier stages wait on it before they reuse a buffer set
  std::mutex mu;
  std::condition_variable cv;
  uint64_t written = 0;
  bool failed = false;
  void done_one() {
    { std::lock_guard<std::mutex> l(mu); written++; }
    cv.notify_all();
  }
  void fail() {
    { std::lock_guard<std::mutex> l(mu); failed = true; }
    cv.notify_all();
  }
  bool wait_written(uint64_t at_least) {  // false when another stage failed
    std::unique_lock<std::mutex> l(mu);
    cv.wait(l, [&] { return failed || written >= at_least; });
    return !failed;
  }
};

}  // namespace

extern "C" tg_status tg_align_files(const tg_index_host* ix, tg_ctx* ctx, tg_multi* multi, const char* const* query_paths,
                                    int n_paths, const char* out_path, int output_fmt, uint32_t batch_reads, tg_file_stats* stats) {
  TG_GUARD_BEGIN
  if (!ix || (!ctx) == (!multi) || !query_paths || n_paths < 1 || !out_path || output_fmt < 0 || output_fmt > 2)
    return tg_fail(TG_ERR_INVALID, "tg_align_files: bad argument (exactly one of ctx / multi; format 0 PAF, 1 SAM, 2 BAM)");
  if (batch_reads == 0) batch_reads = 1u << 20;
  tg_status st;
  if (ctx) st = tg_ctx_set_result_buffers(ctx, 2);
  else st = tg_multi_set_result_buffers(multi, 2);
  if (st != TG_OK) return st;
  FILE* out = strcmp(out_path, "-") == 0 ? stdout : fopen(out_path, "wb");
  if (!out) return tg_fail(TG_ERR_IO, std::string("cannot create ") + out_path);
  auto close_out = [&]() { if (out != stdout) fclose(out); else fflush(out); };
  const double t_start = now_ms();
  tg_file_stats S;
  memset(&S, 0, sizeof(S));
  // file header (src/aligner.rs:41-47)
  if (output_fmt == 1) {
    char* h = nullptr; size_t hl = 0;
    if ((st = tg_format_sam_header(ix, &h, &hl)) != TG_OK) { close_out(); return st; }
    fwrite(h, 1, hl, out); S.bytes_out += hl; free(h);
  } else if (output_fmt == 2) {
    void* h = nullptr; size_t hl = 0;
    if ((st = tg_format_bam_header(ix, &h, &hl)) != TG_OK) { close_out(); return st; }
    fwrite(h, 1, hl, out); S.bytes_out += hl; free(h);
  }
  // a regular file is written with pwrite from many threads; a pipe / terminal through the stdio stream
  fflush(out);
  const int out_fd = fileno(out);
  off_t file_pos = lseek(out_fd, 0, SEEK_CUR);
  bool seekable = false;
  {
    struct stat sb;
    seekable = out != stdout && file_pos >= 0 && fstat(out_fd, &sb) == 0 && S_ISREG(sb.st_mode);  // (stdout may be in append mode)
  }
  // PAF / SAM on one GPU: lines are written on the device (tg_paf.cu), the writers only copy text into the file
  tg_paf* paf = nullptr;
  if (output_fmt <= 1 && ctx && !getenv("TG_PAF_HOST")) {
    st = output_fmt == 0 ? tg_paf_create(ix, ctx, tg_ctx_device(ctx), &paf) : tg_sam_create(ix, ctx, tg_ctx_device(ctx), &paf);
    if (st != TG_OK) { close_out(); return st; }
  }
  Channel to_align, to_write;
  Progress prog;
  std::string err_read, err_write;
  tg_status st_read = TG_OK, st_write = TG_OK;
  double read_ms = 0, write_ms = 0, format_ms = 0;

  // stage 1: files -> batches.  A reader owns three buffer sets; batch b may be overwritten by batch b + 3, which is
  // produced only after batch b has been written.
  std::vector<tg_fastq_reader*> open_readers;
  std::thread producer([&]() {
    lower_priority();
    uint64_t b = 0;
    for (int f = 0; f < n_paths && st_read == TG_OK; f++) {
      tg_fastq_reader* r = nullptr;
      if ((st_read = tg_fastq_open(query_paths[f], &r)) != TG_OK) { err_read = tg_last_error(); break; }
      // (page-locked qualities for SAM on the GPU were tried: 7 ms less copy per 1 M reads in the aligner stage, which is not the
      // bound, against three more page-locked allocations in the reader: 8 M reads went from 5.4 to 4.4 M reads/s overall)
      for (;;) {
        if (b >= 2 && !prog.wait_written(b - 2)) break;
        Job j;
        const double t0 = now_ms();
        if ((st_read = tg_fastq_next(r, batch_reads, &j.batch)) != TG_OK) { err_read = tg_last_error(); break; }
        read_ms += now_ms() - t0;
        if (j.batch.n_reads == 0) break;
        // the device path takes reads of at most TG_MAX_READ_LEN bases (include/thermite_gpu.h, limits): name the read here
        // instead of letting the aligner refuse the whole batch
        for (uint32_t i = 0; i < j.batch.n_reads && st_read == TG_OK; i++) {
          const uint64_t len = j.batch.offs[i + 1] - j.batch.offs[i];
          if (len <= TG_MAX_READ_LEN) continue;
          st_read = TG_ERR_INVALID;
          err_read = "read '" + std::string((const char*)j.batch.names + j.batch.name_offs[i], (size_t)std::min<uint64_t>(j.batch.name_offs[i + 1] - j.batch.name_offs[i], 200)) +
                     "' of " + query_paths[f] + " has " + std::to_string(len) + " bases; reads of at most " + std::to_string(TG_MAX_READ_LEN) +
                     " are aligned (batches before it are already in the output)";
        }
        if (st_read != TG_OK) break;
        j.index = b++;
        to_align.push(j);
      }
      // the reader's buffers must outlive its batches: close it once they are written; after a failure in any stage the
      // batches on their way may still be read, so the reader is closed when the pipeline has come to rest
      if (st_read == TG_OK && prog.wait_written(b)) tg_fastq_close(r);
      else { open_readers.push_back(r); break; }
    }
    if (st_read != TG_OK) prog.fail();
    Job e; e.end = true;
    to_align.push(e);
  });

  // stage 3: records -> text -> file.  Per-thread text buffers are kept across batches (fresh buffers cost a page fault
  // per 4 KiB of output) and written one after the other: no merged copy of the batch's text is ever made.
  std::thread writer([&]() {
    if (!paf) lower_priority();  // (with the text written on the GPU this thread only copies it into the file: the bound of SAM output)
    const unsigned T = host_threads();
    std::vector<TgOut> parts(T);
    std::vector<std::string> zparts;
    for (;;) {
      Job j = to_write.pop();
      if (j.end) break;
      if (st_write != TG_OK) { prog.done_one(); continue; }  // keep draining so that nobody waits forever
      const double t0 = now_ms();
      const uint32_t n = j.batch.n_reads;
      if (paf) {  // the text arrives by itself: wait for the copy, then put it into the file
        if (tg_paf_wait(paf, j.text) != TG_OK) { st_write = TG_ERR_CUDA; err_write = tg_last_error(); }
        // (writes into one file take turns on its lock: one thread is as fast as sixteen and leaves the cores to the parser;
        // TG_WRITE_THREADS sets another number)
        static const unsigned write_threads = []() { const char* e = getenv("TG_WRITE_THREADS"); const long v = e ? atol(e) : 0; return v > 0 ? (unsigned)std::min<long>(v, 64) : 0u; }();
        const unsigned Tw = (!seekable || j.text_len < (8u << 20)) ? 1u : write_threads ? std::min(write_threads, T) : 1u;
        if (st_write != TG_OK) {
          // (the copy failed: nothing of this batch is written)
        } else if (!seekable) {
          if (j.text_len && fwrite(j.text, 1, j.text_len, out) != j.text_len) { st_write = TG_ERR_IO; err_write = "write error"; }
        } else {
          std::vector<int> bad(Tw, 0);
          run_threads(Tw, [&](unsigned t) {
            size_t a = j.text_len * t / Tw, b = j.text_len * (t + 1) / Tw;
            while (a < b) {
              const ssize_t w = pwrite(out_fd, j.text + a, b - a, file_pos + (off_t)a);
              if (w <= 0) { bad[t] = 1; return; }
              a += (size_t)w;
            }
          });
          for (int x : bad) if (x) { st_write = TG_ERR_IO; err_write = "write error"; }
          file_pos += (off_t)j.text_len;
        }
        S.bytes_out += j.text_len;
        S.n_reads += n; S.n_alns += j.res.n_alns; S.n_batches++;
        if (S.n_batches == 2) { S.warm_reads = S.n_reads; S.warm_ms = now_ms() - t_start; }
        write_ms += now_ms() - t0;
        if (st_write != TG_OK) prog.fail();
        prog.done_one();
        continue;
      }
      const unsigned Tn = n < 32768 ? 1u : T;
      TgRecView v;
      v.first32 = j.res.read_aln_first; v.count = j.res.read_aln_count; v.comp = j.res.alns; v.ops = j.res.ops;
      run_threads(Tn, [&](unsigned t) {
        parts[t].s.n = 0;
        const uint32_t r0 = (uint32_t)((uint64_t)n * t / Tn), r1 = (uint32_t)((uint64_t)n * (t + 1) / Tn);
        tg_format_reads(ix, v, j.batch.bases, j.batch.offs, j.batch.names, j.batch.name_offs, j.batch.quals, j.batch.qual_offs,
                        output_fmt == 0 ? 0 : 1, r0, r1, parts[t]);
      });
      format_ms += now_ms() - t0;
      if (output_fmt != 2 && seekable) {
        // every thread copies its own piece into the file (page cache) at its final offset
        std::vector<off_t> at(Tn + 1, file_pos);
        for (unsigned t = 0; t < Tn; t++) at[t + 1] = at[t] + (off_t)parts[t].s.size();
        std::vector<int> bad(Tn, 0);
        run_threads(Tn, [&](unsigned t) {
          const char* p = parts[t].s.data();
          size_t len = parts[t].s.size();
          off_t o = at[t];
          while (len) {
            const ssize_t w = pwrite(out_fd, p, len, o);
            if (w <= 0) { bad[t] = 1; return; }
            p += w; len -= (size_t)w; o += w;
          }
        });
        for (int b : bad) if (b) { st_write = TG_ERR_IO; err_write = "write error"; }
        S.bytes_out += (uint64_t)(at[Tn] - file_pos);
        file_pos = at[Tn];
      } else {
        if (output_fmt == 2) {  // every thread turns its own text into BAM records and BGZF blocks; the pieces go out in order
          if (zparts.size() < Tn) zparts.resize(Tn);
          std::vector<tg_status> zst(Tn, TG_OK);
          std::vector<std::string> zerr(Tn);
          run_threads(Tn, [&](unsigned t) {
            zparts[t].clear();
            zst[t] = tg_sam_text_to_bam(ix, parts[t].s.data(), parts[t].s.size(), false, zparts[t], 1);
            if (zst[t] != TG_OK) zerr[t] = tg_last_error();
          });
          for (unsigned t = 0; t < Tn && st_write == TG_OK; t++) if (zst[t] != TG_OK) { st_write = zst[t]; err_write = zerr[t]; }
        }
        for (unsigned t = 0; t < Tn && st_write == TG_OK; t++) {
          const char* p = output_fmt == 2 ? zparts[t].data() : parts[t].s.data();
          const size_t len = output_fmt == 2 ? zparts[t].size() : parts[t].s.size();
          if (len && fwrite(p, 1, len, out) != len) { st_write = TG_ERR_IO; err_write = "write error"; }
          S.bytes_out += len;
        }
      }
      S.n_reads += n; S.n_alns += j.res.n_alns; S.n_batches++;
      if (S.n_batches == 2) { S.warm_reads = S.n_reads; S.warm_ms = now_ms() - t_start; }  // first batches: buffers grow, a batch may be re-run
      write_ms += now_ms() - t0;
      if (st_write != TG_OK) prog.fail();
      prog.done_one();
    }
  });

  // stage 2 (this thread): batches -> records.  Two result sets alternate: batch b overwrites the result of batch b - 2.
  tg_status st_align = TG_OK;
  double align_ms = 0;
  for (;;) {
    Job j = to_align.pop();
    if (j.end) break;
    if (st_align != TG_OK) continue;
    if (j.index >= 1 && !prog.wait_written(j.index - 1)) { st_align = TG_ERR_INTERNAL; continue; }
    const double t0 = now_ms();
    if (paf) {
      tg_result counters;
      memset(&j.res, 0, sizeof(j.res));
      st_align = tg_paf_align_batch_async(paf, &j.batch, &j.text, &j.text_len, &counters);  // the writer waits for the text
      if (st_align == TG_OK) { j.res.n_reads = j.batch.n_reads; j.res.n_alns = counters.n_alns; j.res.n_ops = counters.n_ops; }
    } else if (ctx) st_align = tg_align_batch_compact(ctx, j.batch.bases, j.batch.offs, j.batch.n_reads, &j.res);
    else st_align = tg_multi_align_batch(multi, j.batch.bases, j.batch.offs, j.batch.n_reads, &j.res);
    align_ms += now_ms() - t0;
    if (st_align != TG_OK) { prog.fail(); continue; }
    to_write.push(j);
  }
  { Job e; e.end = true; to_write.push(e); }
  producer.join();
  writer.join();
  for (tg_fastq_reader* r : open_readers) tg_fastq_close(r);
  if (output_fmt == 2 && st_align == TG_OK && st_read == TG_OK && st_write == TG_OK) {
    static const unsigned char eof_block[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof_block, 1, sizeof(eof_block), out);
    S.bytes_out += sizeof(eof_block);
  }
  if (paf) tg_paf_destroy(paf);
  close_out();
  S.read_ms = read_ms; S.align_ms = align_ms; S.write_ms = write_ms; S.format_ms = format_ms; S.wall_ms = now_ms() - t_start;
  if (stats) *stats = S;
  if (st_read != TG_OK) return tg_fail(st_read, err_read);
  if (st_align != TG_OK) return st_align == TG_ERR_INTERNAL && prog.failed ? tg_fail(st_write != TG_OK ? st_write : TG_ERR_INTERNAL, err_write.empty() ? "pipeline stage failed" : err_write) : st_align;
  if (st_write != TG_OK) return tg_fail(st_write, err_write);
  return TG_OK;
  TG_GUARD_END
}
