// host_batcher.cpp -- per-read align_read calls from many host threads, served by ONE GPU context.
//
// The reference's embedding API is ThermiteAligner (/root/reference/src/wrapper.rs:20-27): `Clone + Send` around an
// `Arc<Index>`, one clone per worker thread, `align_read(name, read, qual)` per read (:72) -> align_read
// (src/aligner.rs:123).  A GPU context wants batches.  tg_batcher sits between the two: callers submit reads (ticket) and
// wait for them -- or block in tg_batcher_align_read, which is submit + wait -- and a dispatcher thread runs one
// tg_align_batch over what is queued as soon as `max_batch_reads` are waiting or the oldest request is `max_wait_us`
// old (so everything that arrived while the previous batch was on the GPU leaves at once).  Every caller gets its own
// records (the Vec<GenomeAlignment> of its read, operations rebased to a private block).  Results are those of
// tg_align_batch, which does not depend on how reads are batched.
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "tg_internal.h"

namespace {

using Clock = std::chrono::steady_clock;

struct Request {
  std::vector<uint8_t> read;  // copied at submit: the caller's buffer is free again when submit returns
  Clock::time_point arrival;
  tg_read_alns res{0, 0, nullptr, nullptr};
  tg_status status = TG_OK;
  bool done = false;
};

void free_alns(tg_read_alns* r) {
  free(r->alns);  // records and operations are one block
  memset(r, 0, sizeof(*r));
}

}  // namespace

struct tg_batcher {
  tg_batch_backend_fn fn = nullptr;
  void* user = nullptr;
  uint32_t max_batch = 0;
  std::chrono::microseconds max_wait{0};

  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<Request*> queue;                        // submitted, not yet on the GPU
  std::unordered_map<uint64_t, Request*> tickets;    // submitted, not yet waited for
  uint64_t next_ticket = 1;
  bool stop = false;
  std::string error;  // message of the last failed batch (copied into the waiters' thread-local error)
  uint64_t n_reads = 0, n_batches = 0;
  uint32_t largest = 0;
  std::thread worker;

  void run();
  void serve(std::vector<Request*>& batch, std::vector<uint8_t>& bases, std::vector<uint64_t>& offs);
};

void tg_batcher::run() {
  std::vector<Request*> batch;
  std::vector<uint8_t> bases;
  std::vector<uint64_t> offs;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_work.wait(lk, [&] { return stop || !queue.empty(); });
      if (queue.empty()) return;  // stop requested and nothing left to serve
      // give concurrent callers until the oldest request is max_wait old to join; a full batch goes at once
      const auto deadline = queue.front()->arrival + max_wait;
      cv_work.wait_until(lk, deadline, [&] { return stop || queue.size() >= max_batch; });
      const size_t take = std::min<size_t>(queue.size(), max_batch);
      batch.assign(queue.begin(), queue.begin() + take);
      queue.erase(queue.begin(), queue.begin() + take);
    }
    serve(batch, bases, offs);
  }
}

void tg_batcher::serve(std::vector<Request*>& batch, std::vector<uint8_t>& bases, std::vector<uint64_t>& offs) {
  const uint32_t n = (uint32_t)batch.size();
  offs.assign(1, 0);
  bases.clear();
  for (Request* r : batch) {
    bases.insert(bases.end(), r->read.begin(), r->read.end());
    offs.push_back(bases.size());
  }
  if (bases.empty()) bases.push_back('N');  // keep data() non-null for a batch of empty reads
  tg_result res;
  memset(&res, 0, sizeof(res));
  tg_status st = TG_ERR_INTERNAL;
  std::string msg;
  try {
    st = fn(user, bases.data(), offs.data(), n, &res);
    if (st != TG_OK) msg = tg_last_error();
  } catch (const std::exception& e) {
    msg = e.what();
  } catch (...) {
    msg = "unknown exception in the batch backend";
  }
  if (st == TG_OK) {
    for (uint32_t i = 0; i < n && st == TG_OK; i++) {
      tg_read_alns* o = &batch[i]->res;
      const uint64_t first = res.read_aln_first[i];
      const uint32_t cnt = res.read_aln_count[i];
      uint64_t n_ops = 0;
      for (uint32_t a = 0; a < cnt; a++) n_ops += (uint64_t)res.alns[first + a].ops_len + res.alns[first + a].tx_ops_len;
      o->n_alns = cnt;
      o->n_ops = (uint32_t)n_ops;
      if (!cnt) continue;
      // one block: records, then their operation words
      char* blk = (char*)malloc(cnt * sizeof(tg_aln) + n_ops * sizeof(uint32_t));
      if (!blk) { st = TG_ERR_INTERNAL; msg = "out of memory"; break; }
      o->alns = (tg_aln*)blk;
      o->ops = (uint32_t*)(blk + cnt * sizeof(tg_aln));
      uint32_t w = 0;
      for (uint32_t a = 0; a < cnt; a++) {
        tg_aln rec = res.alns[first + a];
        if (rec.ops_len) memcpy(o->ops + w, res.ops + rec.ops_off, (size_t)rec.ops_len * 4);
        rec.ops_off = w;
        w += rec.ops_len;
        if (rec.tx_ops_len) memcpy(o->ops + w, res.ops + rec.tx_ops_off, (size_t)rec.tx_ops_len * 4);
        rec.tx_ops_off = rec.tx_ops_len ? w : 0;
        w += rec.tx_ops_len;
        o->alns[a] = rec;
      }
    }
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (st != TG_OK) error = msg;
    n_reads += n;
    n_batches++;
    largest = std::max(largest, n);
    for (Request* r : batch) {
      r->status = st;
      r->done = true;
    }
  }
  cv_done.notify_all();
}

tg_status tg_batcher_create_backend(tg_batch_backend_fn fn, void* user, uint32_t max_batch_reads, uint32_t max_wait_us,
                                    tg_batcher** out) {
  if (!fn || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  if (max_batch_reads == 0) return tg_fail(TG_ERR_INVALID, "max_batch_reads must be at least 1");
  try {
    auto* b = new tg_batcher();
    b->fn = fn;
    b->user = user;
    b->max_batch = max_batch_reads;
    b->max_wait = std::chrono::microseconds(max_wait_us);
    b->worker = std::thread([b] { b->run(); });
    *out = b;
    return TG_OK;
  } catch (const std::exception& e) {
    return tg_fail(TG_ERR_INTERNAL, e.what());
  }
}

extern "C" {

tg_status tg_batcher_submit(tg_batcher* b, const uint8_t* read, uint32_t len, uint64_t* ticket) {
  if (!b || !ticket || (!read && len)) return tg_fail(TG_ERR_INVALID, "null argument");
  if (len > TG_MAX_READ_LEN) return tg_fail(TG_ERR_CAPACITY, "read longer than TG_MAX_READ_LEN");
  Request* rq = nullptr;
  try {
    rq = new Request();
    rq->read.assign(read, read + len);
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->stop) { delete rq; return tg_fail(TG_ERR_INVALID, "the batcher is being destroyed"); }
    rq->arrival = Clock::now();
    *ticket = b->next_ticket++;
    b->tickets.emplace(*ticket, rq);
    b->queue.push_back(rq);
    if (b->queue.size() == 1 || b->queue.size() >= b->max_batch) b->cv_work.notify_one();
    return TG_OK;
  } catch (const std::exception& e) {
    delete rq;
    return tg_fail(TG_ERR_INTERNAL, e.what());
  }
}

tg_status tg_batcher_wait(tg_batcher* b, uint64_t ticket, tg_read_alns* out) {
  if (!b || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  std::unique_lock<std::mutex> lk(b->mu);
  auto it = b->tickets.find(ticket);
  if (it == b->tickets.end()) return tg_fail(TG_ERR_INVALID, "unknown ticket (never issued or already waited for)");
  Request* rq = it->second;
  b->tickets.erase(it);  // a ticket is waited for once
  b->cv_done.wait(lk, [&] { return rq->done; });
  const tg_status st = rq->status;
  const std::string msg = st != TG_OK ? b->error : std::string();
  lk.unlock();
  if (st == TG_OK) *out = rq->res;
  else free_alns(&rq->res);
  delete rq;
  return st == TG_OK ? TG_OK : tg_fail(st, "batched align_read failed: " + msg);
}

tg_status tg_batcher_align_read(tg_batcher* b, const uint8_t* read, uint32_t len, tg_read_alns* out) {
  uint64_t ticket = 0;
  tg_status st = tg_batcher_submit(b, read, len, &ticket);
  return st != TG_OK ? st : tg_batcher_wait(b, ticket, out);
}

void tg_read_alns_free(tg_read_alns* r) {
  if (r) free_alns(r);
}

tg_status tg_batcher_stats(tg_batcher* b, uint64_t* n_reads, uint64_t* n_batches, uint32_t* largest_batch) {
  if (!b) return tg_fail(TG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(b->mu);
  if (n_reads) *n_reads = b->n_reads;
  if (n_batches) *n_batches = b->n_batches;
  if (largest_batch) *largest_batch = b->largest;
  return TG_OK;
}

void tg_batcher_destroy(tg_batcher* b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->stop = true;  // queued requests are still served; new ones are refused
  }
  b->cv_work.notify_all();
  if (b->worker.joinable()) b->worker.join();
  for (auto& kv : b->tickets) {  // results nobody waited for
    free_alns(&kv.second->res);
    delete kv.second;
  }
  delete b;
}

}  // extern "C"
