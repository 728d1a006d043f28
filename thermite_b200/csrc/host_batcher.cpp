// host_batcher.cpp -- per-read align_read calls from many host threads, served by ONE GPU context.
//
// The reference's embedding API is ThermiteAligner (/root/reference/src/wrapper.rs:20-27): `Clone + Send` around an
// `Arc<Index>`, one clone per worker thread, `align_read(name, read, qual)` per read (:72) -> align_read
// (src/aligner.rs:123).  A GPU context wants batches.  tg_batcher sits between the two: callers submit reads (ticket) and
// wait for them -- or block in tg_batcher_align_read, which is submit + wait -- and a dispatcher thread runs one
// tg_align_batch over what is queued as soon as `max_batch_reads` are waiting, the oldest request is `max_wait_us` old, or
// no new request has arrived for max_wait_us / 16 (10..100 us).  Every caller gets its own
// records (the Vec<GenomeAlignment> of its read, operations rebased to a private block).  Results are those of
// tg_align_batch, which does not depend on how reads are batched.
//
// Host cost per read is what limits this path, so: requests live in a slab of reusable slots (no allocation per submit,
// ticket = slot | generation << 32), the dispatcher only snapshots a finished batch (four bulk copies) and marks its
// slots done, and every waiter cuts its own records out of the snapshot on its own thread.
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "tg_internal.h"

namespace {

using Clock = std::chrono::steady_clock;

struct BatchSnapshot {  // host copy of one tg_result (the context's own buffers are reused by the next batch)
  std::vector<uint64_t> first;
  std::vector<uint32_t> count, ops;
  std::vector<tg_aln> alns;
};

struct Slot {
  uint32_t gen = 1;  // bumped when the slot is freed: a stale ticket no longer matches
  uint32_t len = 0;
  bool in_use = false, waited = false, done = false;
  tg_status status = TG_OK;
  Clock::time_point arrival;
  std::shared_ptr<BatchSnapshot> batch;
  uint32_t index = 0;  // the read's row in `batch`
  uint8_t read[TG_MAX_READ_LEN];  // copied at submit: the caller's buffer is free again when submit returns
};

// the records of row i of a snapshot as one malloc'ed block: records, then their operation words (rebased)
tg_status cut_read(const BatchSnapshot& b, uint32_t i, tg_read_alns* o) {
  const uint64_t first = b.first[i];
  const uint32_t cnt = b.count[i];
  uint64_t n_ops = 0;
  for (uint32_t a = 0; a < cnt; a++) n_ops += (uint64_t)b.alns[first + a].ops_len + b.alns[first + a].tx_ops_len;
  o->n_alns = cnt;
  o->n_ops = (uint32_t)n_ops;
  o->alns = nullptr;
  o->ops = nullptr;
  if (!cnt) return TG_OK;
  char* blk = (char*)malloc(cnt * sizeof(tg_aln) + n_ops * sizeof(uint32_t));
  if (!blk) return tg_fail(TG_ERR_INTERNAL, "out of memory");
  o->alns = (tg_aln*)blk;
  o->ops = (uint32_t*)(blk + cnt * sizeof(tg_aln));
  uint32_t w = 0;
  for (uint32_t a = 0; a < cnt; a++) {
    tg_aln rec = b.alns[first + a];
    if (rec.ops_len) memcpy(o->ops + w, b.ops.data() + rec.ops_off, (size_t)rec.ops_len * 4);
    rec.ops_off = w;
    w += rec.ops_len;
    if (rec.tx_ops_len) memcpy(o->ops + w, b.ops.data() + rec.tx_ops_off, (size_t)rec.tx_ops_len * 4);
    rec.tx_ops_off = rec.tx_ops_len ? w : 0;
    w += rec.tx_ops_len;
    o->alns[a] = rec;
  }
  return TG_OK;
}

}  // namespace

struct tg_batcher {
  tg_batch_backend_fn fn = nullptr;
  void* user = nullptr;
  uint32_t max_batch = 0;
  std::chrono::microseconds max_wait{0}, quiet{0};
  Clock::time_point last_arrival;

  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<Slot> slots;          // stable addresses; indexed only under `mu`
  std::vector<uint32_t> free_slots;
  std::deque<Slot*> queue;         // submitted, not yet on the GPU
  bool stop = false;
  std::string error;  // message of the last failed batch (copied into the waiters' thread-local error)
  uint64_t n_reads = 0, n_batches = 0;
  uint32_t largest = 0;
  std::thread worker;

  void run();
  void serve(std::vector<Slot*>& batch, std::vector<uint8_t>& bases, std::vector<uint64_t>& offs);
};

void tg_batcher::run() {
  std::vector<Slot*> batch;
  std::vector<uint8_t> bases;
  std::vector<uint64_t> offs;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_work.wait(lk, [&] { return stop || !queue.empty(); });
      if (queue.empty()) return;  // stop requested and nothing left to serve
      // A batch leaves when it is full, when its oldest request is max_wait old, or when nothing new has arrived for
      // `quiet` (blocking callers all resubmit within microseconds of the previous batch and then wait: no point in
      // sitting out max_wait; callers streaming windows of tickets keep arriving, and the batch keeps growing).
      while (!stop && queue.size() < max_batch) {
        const auto deadline = std::min(queue.front()->arrival + max_wait, last_arrival + quiet);
        if (Clock::now() >= deadline) break;
        cv_work.wait_until(lk, deadline);
      }
      const size_t take = std::min<size_t>(queue.size(), max_batch);
      batch.assign(queue.begin(), queue.begin() + take);
      queue.erase(queue.begin(), queue.begin() + take);
    }
    serve(batch, bases, offs);
  }
}

void tg_batcher::serve(std::vector<Slot*>& batch, std::vector<uint8_t>& bases, std::vector<uint64_t>& offs) {
  const uint32_t n = (uint32_t)batch.size();
  offs.assign(1, 0);
  bases.clear();
  for (Slot* s : batch) {  // (a queued slot is not touched by anyone else until it is marked done)
    bases.insert(bases.end(), s->read, s->read + s->len);
    offs.push_back(bases.size());
  }
  if (bases.empty()) bases.push_back('N');  // keep data() non-null for a batch of empty reads
  tg_result res;
  memset(&res, 0, sizeof(res));
  tg_status st = TG_ERR_INTERNAL;
  std::string msg;
  std::shared_ptr<BatchSnapshot> snap;
  try {
    st = fn(user, bases.data(), offs.data(), n, &res);
    if (st != TG_OK) msg = tg_last_error();
    else {
      snap = std::make_shared<BatchSnapshot>();
      snap->first.assign(res.read_aln_first, res.read_aln_first + n);
      snap->count.assign(res.read_aln_count, res.read_aln_count + n);
      snap->alns.assign(res.alns, res.alns + res.n_alns);
      snap->ops.assign(res.ops, res.ops + res.n_ops);
    }
  } catch (const std::exception& e) {
    st = TG_ERR_INTERNAL;
    msg = e.what();
  } catch (...) {
    st = TG_ERR_INTERNAL;
    msg = "unknown exception in the batch backend";
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (st != TG_OK) error = msg;
    n_reads += n;
    n_batches++;
    largest = std::max(largest, n);
    for (uint32_t i = 0; i < n; i++) {
      Slot* s = batch[i];
      s->status = st;
      s->batch = snap;
      s->index = i;
      s->done = true;
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
    b->quiet = std::chrono::microseconds(std::min<uint32_t>(std::max<uint32_t>(max_wait_us / 16, 10u), 100u));
    if (b->quiet > b->max_wait) b->quiet = b->max_wait;
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
  try {
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->stop) return tg_fail(TG_ERR_INVALID, "the batcher is being destroyed");
    uint32_t idx;
    if (!b->free_slots.empty()) {
      idx = b->free_slots.back();
      b->free_slots.pop_back();
    } else {
      idx = (uint32_t)b->slots.size();
      b->slots.emplace_back();
    }
    Slot& s = b->slots[idx];
    s.in_use = true; s.waited = false; s.done = false; s.status = TG_OK; s.len = len;
    if (len) memcpy(s.read, read, len);
    s.arrival = Clock::now();
    b->last_arrival = s.arrival;
    *ticket = (uint64_t)idx | ((uint64_t)s.gen << 32);
    b->queue.push_back(&s);
    if (b->queue.size() == 1 || b->queue.size() == b->max_batch) b->cv_work.notify_one();
    return TG_OK;
  } catch (const std::exception& e) {
    return tg_fail(TG_ERR_INTERNAL, e.what());
  }
}

tg_status tg_batcher_wait(tg_batcher* b, uint64_t ticket, tg_read_alns* out) {
  if (!b || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  const uint32_t idx = (uint32_t)(ticket & 0xFFFFFFFFull), gen = (uint32_t)(ticket >> 32);
  std::shared_ptr<BatchSnapshot> snap;
  uint32_t row = 0;
  tg_status st;
  std::string msg;
  {
    std::unique_lock<std::mutex> lk(b->mu);
    if (idx >= b->slots.size()) return tg_fail(TG_ERR_INVALID, "unknown ticket (never issued or already waited for)");
    Slot* s = &b->slots[idx];
    if (!s->in_use || s->gen != gen || s->waited) return tg_fail(TG_ERR_INVALID, "unknown ticket (never issued or already waited for)");
    s->waited = true;  // a ticket is waited for once
    b->cv_done.wait(lk, [&] { return s->done; });
    st = s->status;
    if (st != TG_OK) msg = b->error;
    snap = std::move(s->batch);
    row = s->index;
    s->batch.reset();
    s->in_use = false;
    s->gen++;
    if (s->gen == 0) s->gen = 1;
    b->free_slots.push_back(idx);
  }
  if (st != TG_OK) return tg_fail(st, "batched align_read failed: " + msg);
  return cut_read(*snap, row, out);  // on the caller's thread, outside the lock
}

tg_status tg_batcher_align_read(tg_batcher* b, const uint8_t* read, uint32_t len, tg_read_alns* out) {
  uint64_t ticket = 0;
  tg_status st = tg_batcher_submit(b, read, len, &ticket);
  return st != TG_OK ? st : tg_batcher_wait(b, ticket, out);
}

void tg_read_alns_free(tg_read_alns* r) {
  if (!r) return;
  free(r->alns);  // records and operations are one block
  memset(r, 0, sizeof(*r));
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
  delete b;  // results nobody waited for go with their snapshots
}

}  // extern "C"
