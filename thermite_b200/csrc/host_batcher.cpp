// host_batcher.cpp -- per-read align_read calls from many host threads, served by ONE GPU context.
//
// The reference's embedding API is ThermiteAligner (/root/reference/src/wrapper.rs:20-27): `Clone + Send` around an
// `Arc<Index>`, one clone per worker thread, `align_read(name, read, qual)` per read (:72) -> align_read
// (src/aligner.rs:123).  A GPU context wants batches.  tg_batcher sits between the two: callers submit reads (ticket) and
// wait for them -- or block in tg_batcher_align_read, which is submit + wait -- and a dispatcher thread runs one
// tg_align_batch over what is queued as soon as `max_batch_reads` are waiting, the oldest request is `max_wait_us` old, or
// no new request has arrived for max_wait_us / 4 (20..100 us).  Every caller gets its own
// records (the Vec<GenomeAlignment> of its read, operations rebased to a private block).  Results are those of
// tg_align_batch, which does not depend on how reads are batched.
//
// Host cost per read is what limits this path, so: requests live in slabs of reusable slots (no allocation per submit,
// ticket = slot | shard << 24 | generation << 32), a caller thread only ever takes the lock of its own shard (32 shards,
// assigned round-robin per thread), the dispatcher only snapshots a finished batch (four bulk copies) and marks its
// slots done, and every waiter cuts its own records out of the snapshot on its own thread.
#include <atomic>
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

struct Shard {
  std::mutex mu;
  std::condition_variable cv_done;
  std::deque<Slot> slots;          // stable addresses; indexed only under `mu`
  std::vector<uint32_t> free_slots;
  std::vector<Slot*> queue;        // submitted, not yet on the GPU
};

struct tg_batcher {
  static constexpr uint32_t NSHARD = 32, SLOT_BITS = 24;
  tg_batch_backend_fn fn = nullptr;
  void* user = nullptr;
  uint32_t max_batch = 0;
  int64_t max_wait_ns = 0, quiet_ns = 0;

  Shard shards[NSHARD];
  std::atomic<uint32_t> next_shard{0};
  std::atomic<uint64_t> queued{0};                         // requests in the shard queues
  std::atomic<int64_t> oldest_ns{0}, last_arrival_ns{0};   // arrival of the (approximately) oldest / the newest queued request
  std::atomic<bool> stop{false};

  std::mutex mu;  // dispatcher sleep / statistics / error text
  std::condition_variable cv_work;
  std::string error;  // message of the last failed batch (copied into the waiters' thread-local error)
  uint64_t n_reads = 0, n_batches = 0;
  uint32_t largest = 0;
  std::thread worker;

  static int64_t now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now().time_since_epoch()).count(); }
  Shard& my_shard(uint32_t& id) {
    static thread_local uint32_t mine = 0xFFFFFFFFu;
    if (mine == 0xFFFFFFFFu) mine = next_shard.fetch_add(1) % NSHARD;
    id = mine;
    return shards[mine];
  }
  void run();
  void serve(std::vector<Slot*>& batch, const std::vector<uint32_t>& shard_end, std::vector<uint8_t>& bases, std::vector<uint64_t>& offs);
};

void tg_batcher::run() {
  std::vector<Slot*> batch;
  std::vector<uint32_t> shard_end(NSHARD);
  std::vector<uint8_t> bases;
  std::vector<uint64_t> offs;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_work.wait(lk, [&] { return stop.load() || queued.load() > 0; });
      if (queued.load() == 0) return;  // stop requested and nothing left to serve
      // A batch leaves when it is full, when its oldest request is max_wait old, or when nothing new has arrived for
      // `quiet` (blocking callers all resubmit within microseconds of the previous batch and then wait: no point in
      // sitting out max_wait; callers streaming windows of tickets keep arriving, and the batch keeps growing).
      while (!stop.load() && queued.load() < max_batch) {
        const int64_t deadline = std::min(oldest_ns.load() + max_wait_ns, last_arrival_ns.load() + quiet_ns);
        const int64_t now = now_ns();
        if (now >= deadline) break;
        cv_work.wait_for(lk, std::chrono::nanoseconds(deadline - now));
      }
    }
    batch.clear();
    for (uint32_t k = 0; k < NSHARD; k++) {
      Shard& sh = shards[k];
      std::lock_guard<std::mutex> lk(sh.mu);
      const size_t room = max_batch - batch.size();
      const size_t take = std::min(sh.queue.size(), room);
      batch.insert(batch.end(), sh.queue.begin(), sh.queue.begin() + take);
      sh.queue.erase(sh.queue.begin(), sh.queue.begin() + take);
      shard_end[k] = (uint32_t)batch.size();
    }
    if (queued.fetch_sub(batch.size()) > batch.size()) oldest_ns.store(now_ns());  // what stays behind starts a new batch
    if (!batch.empty()) serve(batch, shard_end, bases, offs);
  }
}

void tg_batcher::serve(std::vector<Slot*>& batch, const std::vector<uint32_t>& shard_end, std::vector<uint8_t>& bases,
                       std::vector<uint64_t>& offs) {
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
  }
  uint32_t i = 0;
  for (uint32_t k = 0; k < NSHARD; k++) {
    if (shard_end[k] == i) continue;
    Shard& sh = shards[k];
    {
      std::lock_guard<std::mutex> lk(sh.mu);
      for (; i < shard_end[k]; i++) {
        Slot* s = batch[i];
        s->status = st;
        s->batch = snap;
        s->index = i;
        s->done = true;
      }
    }
    sh.cv_done.notify_all();
  }
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
    b->max_wait_ns = (int64_t)max_wait_us * 1000;
    b->quiet_ns = std::min<int64_t>((int64_t)std::min<uint32_t>(std::max<uint32_t>(max_wait_us / 4, 20u), 100u) * 1000, b->max_wait_ns);
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
  if (b->stop.load()) return tg_fail(TG_ERR_INVALID, "the batcher is being destroyed");
  try {
    uint32_t sid;
    Shard& sh = b->my_shard(sid);
    const int64_t now = tg_batcher::now_ns();
    uint64_t q = 0;
    {
      std::lock_guard<std::mutex> lk(sh.mu);
      uint32_t idx;
      if (!sh.free_slots.empty()) {
        idx = sh.free_slots.back();
        sh.free_slots.pop_back();
      } else {
        idx = (uint32_t)sh.slots.size();
        if (idx >= (1u << tg_batcher::SLOT_BITS)) return tg_fail(TG_ERR_CAPACITY, "too many reads in flight on one thread");
        sh.slots.emplace_back();
      }
      Slot& s = sh.slots[idx];
      s.in_use = true; s.waited = false; s.done = false; s.status = TG_OK; s.len = len;
      if (len) memcpy(s.read, read, len);
      *ticket = (uint64_t)idx | ((uint64_t)sid << tg_batcher::SLOT_BITS) | ((uint64_t)s.gen << 32);
      // counted BEFORE the slot becomes visible to the dispatcher (which drains under this lock and then subtracts what it
      // took): `queued` is never smaller than what sits in the shard queues, so it cannot wrap and queued == 0 means empty
      b->last_arrival_ns.store(now);
      q = b->queued.fetch_add(1) + 1;
      if (q == 1) b->oldest_ns.store(now);
      sh.queue.push_back(&s);
    }
    if (q == 1 || q == b->max_batch) {
      std::lock_guard<std::mutex> lk(b->mu);  // (the dispatcher checks `queued` under this lock before it sleeps)
      b->cv_work.notify_one();
    }
    return TG_OK;
  } catch (const std::exception& e) {
    return tg_fail(TG_ERR_INTERNAL, e.what());
  }
}

tg_status tg_batcher_wait(tg_batcher* b, uint64_t ticket, tg_read_alns* out) {
  if (!b || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  const uint32_t low = (uint32_t)(ticket & 0xFFFFFFFFull), gen = (uint32_t)(ticket >> 32);
  const uint32_t idx = low & ((1u << tg_batcher::SLOT_BITS) - 1), sid = low >> tg_batcher::SLOT_BITS;
  if (sid >= tg_batcher::NSHARD) return tg_fail(TG_ERR_INVALID, "unknown ticket (never issued or already waited for)");
  Shard& sh = b->shards[sid];
  std::shared_ptr<BatchSnapshot> snap;
  uint32_t row = 0;
  tg_status st;
  {
    std::unique_lock<std::mutex> lk(sh.mu);
    if (idx >= sh.slots.size()) return tg_fail(TG_ERR_INVALID, "unknown ticket (never issued or already waited for)");
    Slot* s = &sh.slots[idx];
    if (!s->in_use || s->gen != gen || s->waited) return tg_fail(TG_ERR_INVALID, "unknown ticket (never issued or already waited for)");
    s->waited = true;  // a ticket is waited for once
    sh.cv_done.wait(lk, [&] { return s->done; });
    st = s->status;
    snap = std::move(s->batch);
    row = s->index;
    s->batch.reset();
    s->in_use = false;
    s->gen++;
    if (s->gen == 0) s->gen = 1;
    sh.free_slots.push_back(idx);
  }
  if (st != TG_OK) {
    std::string msg;
    {
      std::lock_guard<std::mutex> lk(b->mu);
      msg = b->error;
    }
    return tg_fail(st, "batched align_read failed: " + msg);
  }
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
    b->stop.store(true);  // queued requests are still served; new ones are refused
  }
  b->cv_work.notify_all();
  if (b->worker.joinable()) b->worker.join();
  delete b;  // results nobody waited for go with their snapshots
}

}  // extern "C"
