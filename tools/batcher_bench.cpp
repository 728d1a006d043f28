// batcher_bench.cpp -- throughput of the per-read API (tg_batcher_*) driven the way the reference's
// ThermiteAligner is (src/wrapper.rs:20-27): T worker threads, each aligning its own share of the reads one read per call,
// either blocking (window 1) or with `window` reads in flight (submit / wait).  Native threads, so the number is the
// library's and not an interpreter's.  Built by tools/batcher_bench.py against libthermite_gpu.so.
// usage: batcher_bench <index.tai> <reads.bin> <threads> <window> <max_batch_reads> <max_wait_us> [reads to use]  (flags -k20 -s0 --intron-mode)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>
#include <vector>

#include "../include/thermite_gpu.h"

#define CK(x)                                                                 \
  do {                                                                        \
    if ((x) != TG_OK) {                                                       \
      fprintf(stderr, "%s failed: %s\n", #x, tg_last_error());                \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s index reads threads window max_batch max_wait_us\n", argv[0]); return 2; }
  const int T = atoi(argv[3]), W = atoi(argv[4]);
  const uint32_t max_batch = (uint32_t)atol(argv[5]), max_wait = (uint32_t)atol(argv[6]);
  // reads.bin: u64 n, u64 offs[n+1], bases
  std::ifstream f(argv[2], std::ios::binary);
  uint64_t n = 0;
  f.read((char*)&n, 8);
  std::vector<uint64_t> offs(n + 1);
  f.read((char*)offs.data(), (n + 1) * 8);
  std::vector<uint8_t> bases(offs[n]);
  f.read((char*)bases.data(), offs[n]);
  if (!f) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
  if (argc > 7 && (uint64_t)atol(argv[7]) < n) n = (uint64_t)atol(argv[7]);

  tg_index_host* hix = nullptr;
  CK(tg_index_host_load(argv[1], &hix));
  tg_index* ix = nullptr;
  CK(tg_index_create(hix, 0, &ix));
  tg_opts o;
  tg_opts_default(&o);
  o.min_seed_len = 20; o.min_aln_score_percent = 0.0f; o.min_aln_score = 30; o.multimap_score_range = 1; o.intron_mode = 1;
  tg_ctx* ctx = nullptr;
  CK(tg_ctx_create(ix, &o, &ctx));
  tg_result res;
  CK(tg_align_batch(ctx, bases.data(), offs.data(), (uint32_t)n, &res));  // warm-up + the numbers to compare with
  const uint64_t want_alns = res.n_alns, want_ops = res.n_ops;
  auto t0 = std::chrono::steady_clock::now();
  CK(tg_align_batch(ctx, bases.data(), offs.data(), (uint32_t)n, &res));
  const double batch_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  tg_batcher* b = nullptr;
  CK(tg_batcher_create(ctx, max_batch, max_wait, &b));
  std::vector<uint64_t> alns(T, 0), ops(T, 0);
  t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t] {
      const uint64_t lo = n * t / T, hi = n * (t + 1) / T;
      std::vector<uint64_t> tk(W);
      tg_read_alns ra;
      for (uint64_t r = lo; r < hi; r += W) {
        const uint64_t m = std::min<uint64_t>(W, hi - r);
        if (W == 1) {
          CK(tg_batcher_align_read(b, bases.data() + offs[r], (uint32_t)(offs[r + 1] - offs[r]), &ra));
          alns[t] += ra.n_alns; ops[t] += ra.n_ops;
          tg_read_alns_free(&ra);
          continue;
        }
        for (uint64_t k = 0; k < m; k++)
          CK(tg_batcher_submit(b, bases.data() + offs[r + k], (uint32_t)(offs[r + k + 1] - offs[r + k]), &tk[k]));
        for (uint64_t k = 0; k < m; k++) {
          CK(tg_batcher_wait(b, tk[k], &ra));
          alns[t] += ra.n_alns; ops[t] += ra.n_ops;
          tg_read_alns_free(&ra);
        }
      }
    });
  for (auto& x : th) x.join();
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t reads = 0, batches = 0, ta = 0, to = 0;
  uint32_t largest = 0;
  CK(tg_batcher_stats(b, &reads, &batches, &largest));
  for (int t = 0; t < T; t++) { ta += alns[t]; to += ops[t]; }
  printf("threads %3d window %6d max_batch %7u max_wait %4u us: %9.0f reads/s per-read API (%llu reads, %llu batches, mean %.0f, "
         "largest %u); one tg_align_batch of all reads: %.0f reads/s; records %s\n",
         T, W, max_batch, max_wait, n / s, (unsigned long long)reads, (unsigned long long)batches, (double)reads / batches, largest,
         n / batch_s, (ta == want_alns && to == want_ops) ? "match the batch call (counts)" : "MISMATCH");
  tg_batcher_destroy(b);
  tg_ctx_destroy(ctx);
  tg_index_destroy(ix);
  tg_index_host_destroy(hix);
  return (ta == want_alns && to == want_ops) ? 0 : 1;
}
