// thermite_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see thermite_oracle.hpp header).
// Every function cites the reference lines it restates.  Parity of the pieces that live in
// un-vendored crates (rust-bio, divsufsort, transcriptome, needletail, noodles) is UNPINNED.
#include "thermite_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <mutex>
#include <numeric>
#include <sstream>
#include <unordered_map>

namespace orc {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
static inline size_t sat_sub(size_t a, size_t b) { return a > b ? a - b : 0; }

// bio::alphabets::dna::complement restricted to the symbols this path can see
static inline uint8_t complement(uint8_t c) {
  switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'a': return 't';
    case 't': return 'a';
    case 'c': return 'g';
    case 'g': return 'c';
    default: return c;  // N -> N, $ -> $
  }
}

std::vector<uint8_t> revcomp(const uint8_t* s, size_t n) {
  std::vector<uint8_t> r(n);
  for (size_t i = 0; i < n; i++) r[i] = complement(s[n - 1 - i]);
  return r;
}

static inline uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

static std::vector<std::string> split_lines(const std::string& text) {
  std::vector<std::string> lines;
  size_t p = 0;
  while (p < text.size()) {
    size_t e = text.find('\n', p);
    if (e == std::string::npos) e = text.size();
    size_t ee = e;
    if (ee > p && text[ee - 1] == '\r') ee--;
    lines.emplace_back(text.substr(p, ee - p));
    p = e + 1;
  }
  return lines;
}

// needletail::parse_fastx_file for FASTA (multi-line joined, id = header without '>')
std::vector<FastaRecord> parse_fasta(const std::string& text) {
  std::vector<FastaRecord> out;
  for (auto& ln : split_lines(text)) {
    if (!ln.empty() && ln[0] == '>') {
      out.push_back(FastaRecord{ln.substr(1), {}});
    } else if (!out.empty()) {
      out.back().seq.insert(out.back().seq.end(), ln.begin(), ln.end());
    }
  }
  return out;
}

// needletail::parse_fastx_file for FASTQ (4-line records)
std::vector<FastqRecord> parse_fastq(const std::string& text) {
  std::vector<FastqRecord> out;
  auto lines = split_lines(text);
  size_t i = 0;
  while (i < lines.size()) {
    if (lines[i].empty()) { i++; continue; }
    if (i + 3 >= lines.size()) break;  // truncated record
    FastqRecord r;
    r.id = lines[i].substr(1);
    r.seq.assign(lines[i + 1].begin(), lines[i + 1].end());
    r.qual.assign(lines[i + 3].begin(), lines[i + 3].end());
    out.push_back(std::move(r));
    i += 4;
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// SwgExtend  (src/swg.rs)
// ---------------------------------------------------------------------------------------------
// src/swg.rs:17-26
SwgExtend::SwgExtend(size_t mbw, int gap_open, int gap_extend, int match, int mismatch)
    : D(mbw * 2 + 1, 0), C(mbw * 2 + 1, 0), R(mbw * 2 + 1, 0), go(gap_open), ge(gap_extend),
      ma(match), mi(mismatch), max_band_width(mbw) {}

// src/swg.rs:210-217 -- grows by ONE row only; an index past the end is a Rust panic (quirk Q4)
void SwgExtend::set_trace(size_t j, size_t i, uint8_t op) {
  size_t w = max_band_width * 2 + 1;
  if (trace.size() <= j * w) trace.resize(trace.size() + w, (uint8_t)Match);
  if (j * w + i >= trace.size())
    throw ReferencePanic("swg.rs:216 index out of bounds (phase-1 x-drop break, x_drop < band)");
  trace[j * w + i] = op;
}
// src/swg.rs:220-223
uint8_t SwgExtend::get_trace(size_t j, size_t i) const {
  size_t w = max_band_width * 2 + 1;
  if (j * w + i >= trace.size()) throw ReferencePanic("swg.rs:222 index out of bounds");
  return trace[j * w + i];
}

// src/swg.rs:226-240  (tie priority diag > Del > Ins)
static inline void triple_max(int32_t d, int32_t c, int32_t r, bool m, int32_t& score, uint8_t& dir) {
  score = std::max(std::max(d, c), r);
  if (score == d) dir = m ? Match : Subst;
  else if (score == c) dir = Del;
  else dir = Ins;
}

// src/swg.rs:31-167
Alignment SwgExtend::extend(const uint8_t* x, size_t xlen, const uint8_t* y, size_t ylen,
                            size_t band_width, int32_t x_drop) {
  if (band_width > max_band_width)  // :32-37 assert!
    throw ReferencePanic("swg.rs:32 band width exceeds max band width");
  Alignment res;
  res.ylen = ylen;
  res.xlen = xlen;
  if (xlen == 0 || ylen == 0) {  // :39-55
    if (xlen > 0) res.operations.push_back(Op{Xclip, xlen});
    return res;
  }
  const size_t w = band_width * 2 + 1;  // :57
  int32_t max_score = 0;
  size_t max_i = 0, max_j = 0;

  // :61-71 leftmost column
  D[0] = 0; C[0] = 0; R[0] = 0;
  set_trace(0, 0, Ins);
  for (size_t i = 1; i < w; i++) {
    C[i] = MIN_SCORE;
    R[i] = (int32_t)i * ge + go;
    D[i] = R[i];
    set_trace(0, i, Ins);
  }

  // :75-113 columns whose band starts at row 0
  for (size_t j = 1; j <= std::min(band_width, ylen); j++) {
    int32_t band_max = MIN_SCORE, prev_D = MIN_SCORE;
    for (size_t i = 0; i < std::min(w, xlen + 1); i++) {
      cells++;
      C[i] = std::max(C[i] + ge, D[i] + ge + go);
      R[i] = (i == 0) ? MIN_SCORE : std::max(R[i - 1] + ge, D[i - 1] + ge + go);
      int32_t d = (i == 0) ? MIN_SCORE : prev_D + (x[i - 1] == y[j - 1] ? ma : mi);
      prev_D = D[i];
      int32_t cur; uint8_t dir;
      triple_max(d, C[i], R[i], i > 0 && x[i - 1] == y[j - 1], cur, dir);
      D[i] = cur;
      set_trace(j, i, dir);
      if (D[i] > max_score) { max_score = D[i]; max_i = i; max_j = j; }
      band_max = std::max(band_max, D[i]);
    }
    if (band_max < max_score - x_drop) break;  // :110-112 (does NOT skip the second loop: Q4)
  }

  // :116-154 columns whose band shifts down by one row per column
  for (size_t j = band_width + 1; j < ylen + 1; j++) {
    int32_t band_max = MIN_SCORE;
    size_t lo = j - band_width, hi = std::min(j - band_width + w, xlen + 1);
    for (size_t i = lo; i < hi; i++) {
      cells++;
      size_t b = i - (j - band_width);
      C[b] = (b >= w - 1) ? MIN_SCORE : std::max(C[b + 1] + ge, D[b + 1] + ge + go);
      R[b] = (b == 0) ? MIN_SCORE : std::max(R[b - 1] + ge, D[b - 1] + ge + go);
      int32_t d = D[b] + (x[i - 1] == y[j - 1] ? ma : mi);
      int32_t cur; uint8_t dir;
      triple_max(d, C[b], R[b], x[i - 1] == y[j - 1], cur, dir);
      D[b] = cur;
      set_trace(j, b, dir);
      if (D[b] > max_score) { max_score = D[b]; max_i = i; max_j = j; }
      band_max = std::max(band_max, D[b]);
    }
    if (band_max < max_score - x_drop) break;  // :150-153
  }

  res.score = max_score;
  res.yend = max_j;
  res.xend = max_i;
  res.operations = trace_path(max_i, max_j, xlen, band_width);
  return res;
}

// src/swg.rs:170-207
std::vector<Op> SwgExtend::trace_path(size_t i, size_t j, size_t len, size_t band_width) const {
  std::vector<Op> tb;
  if (i < len) tb.push_back(Op{Xclip, len - i});
  while (i > 0 || j > 0) {
    size_t b = i - sat_sub(j, band_width);
    uint8_t op = get_trace(j, b);
    tb.push_back(Op{op, 1});
    switch (op) {
      case Match: case Subst: i--; j--; break;
      case Ins: i--; break;
      case Del: j--; break;
      default: throw ReferencePanic("swg.rs:201 unreachable");
    }
  }
  std::reverse(tb.begin(), tb.end());
  return tb;
}

// ---------------------------------------------------------------------------------------------
// IntervalTree (rust-bio AVL)  [recalled, unpinned]
// ---------------------------------------------------------------------------------------------
using Node = IntervalTree::Node;
static int h(const std::unique_ptr<Node>& n) { return n ? n->height : 0; }
static void update(Node* n) {
  n->height = 1 + std::max(h(n->left), h(n->right));
  n->max = n->end;
  if (n->left && n->left->max > n->max) n->max = n->left->max;
  if (n->right && n->right->max > n->max) n->max = n->right->max;
}
static void rotate_left(std::unique_ptr<Node>& n) {
  std::unique_ptr<Node> r = std::move(n->right);
  n->right = std::move(r->left);
  update(n.get());
  r->left = std::move(n);
  n = std::move(r);
  update(n.get());
}
static void rotate_right(std::unique_ptr<Node>& n) {
  std::unique_ptr<Node> l = std::move(n->left);
  n->left = std::move(l->right);
  update(n.get());
  l->right = std::move(n);
  n = std::move(l);
  update(n.get());
}
static void repair(std::unique_ptr<Node>& n) {
  int lh = h(n->left), rh = h(n->right);
  if (std::abs(lh - rh) <= 1) {
    update(n.get());
  } else if (rh > lh) {
    if (h(n->right->left) > h(n->right->right)) rotate_right(n->right);
    rotate_left(n);
  } else {
    if (h(n->left->right) > h(n->left->left)) rotate_left(n->left);
    rotate_right(n);
  }
}
static void insert_node(std::unique_ptr<Node>& n, size_t s, size_t e, size_t d) {
  if (!n) {
    n.reset(new Node{s, e, d, e, 1, nullptr, nullptr});
    return;
  }
  if (s <= n->start) insert_node(n->left, s, e, d);  // ties descend LEFT
  else insert_node(n->right, s, e, d);
  repair(n);
}
void IntervalTree::insert(size_t start, size_t end, size_t data) { insert_node(root, start, end, data); }

std::vector<size_t> IntervalTree::find(size_t qs, size_t qe) const {
  std::vector<size_t> out;
  std::vector<const Node*> stack;
  if (root) stack.push_back(root.get());
  while (!stack.empty()) {
    const Node* c = stack.back();
    stack.pop_back();
    if (qs < c->max) {
      if (c->left) stack.push_back(c->left.get());
      if (qe > c->start) {
        if (c->right) stack.push_back(c->right.get());
        if (c->start < qe && qs < c->end) out.push_back(c->data);
      }
    }
  }
  return out;
}
std::vector<IntervalTree::Flat> IntervalTree::flatten() const {
  std::vector<Flat> out;
  std::vector<const Node*> stack;
  if (root) stack.push_back(root.get());
  while (!stack.empty()) {
    const Node* c = stack.back();
    stack.pop_back();
    if (c->left) stack.push_back(c->left.get());
    if (c->right) stack.push_back(c->right.get());
    out.push_back(Flat{c->start, c->end, c->data, out.size()});
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// suffix array: prefix doubling (Larsson-Sadakane flavour).  Plain byte order; a suffix that is a
// proper prefix of another sorts first (divsufsort64 semantics, src/index.rs:103-105).
// ---------------------------------------------------------------------------------------------
std::vector<uint32_t> suffix_array(const std::vector<uint8_t>& text) {
  const size_t n = text.size();
  std::vector<uint32_t> sa(n);
  if (n == 0) return sa;
  // dense codes 1..d, 0 = past the end
  int code[256];
  std::fill(code, code + 256, -1);
  {
    bool seen[256] = {false};
    for (uint8_t c : text) seen[c] = true;
    int d = 0;
    for (int c = 0; c < 256; c++) if (seen[c]) code[c] = ++d;
  }
  int d = 0;
  for (int c = 0; c < 256; c++) d = std::max(d, code[c]);
  const uint64_t base = (uint64_t)d + 1;
  size_t h0 = 1;
  uint64_t nb = base;
  while (nb * base <= (1ull << 25)) { nb *= base; h0++; }
  // key of the h0-mer at every position (rolling from the right)
  std::vector<uint32_t> key(n);
  {
    uint64_t k = 0, top = nb / base;
    for (size_t t = 0; t < h0; t++) k = k * base;  // all past-the-end
    for (size_t i = n; i-- > 0;) {
      k = k / base + (uint64_t)code[text[i]] * top;
      key[i] = (uint32_t)k;
    }
  }
  std::vector<uint32_t> cnt(nb + 1, 0);
  for (size_t i = 0; i < n; i++) cnt[key[i] + 1]++;
  for (size_t b = 0; b < nb; b++) cnt[b + 1] += cnt[b];
  {
    std::vector<uint32_t> pos(cnt.begin(), cnt.end() - 1);
    for (size_t i = 0; i < n; i++) sa[pos[key[i]]++] = (uint32_t)i;
  }
  // rank = index of the first element of the suffix's group
  std::vector<uint32_t> rnk(n);
  std::vector<std::pair<uint32_t, uint32_t>> groups, next_groups;  // unsorted [s,e)
  {
    size_t s = 0;
    while (s < n) {
      size_t e = s + 1;
      uint32_t k = key[sa[s]];
      while (e < n && key[sa[e]] == k) e++;
      for (size_t i = s; i < e; i++) rnk[sa[i]] = (uint32_t)s;
      if (e - s > 1) groups.emplace_back((uint32_t)s, (uint32_t)e);
      s = e;
    }
  }
  std::vector<uint32_t>().swap(key);
  std::vector<uint32_t>().swap(cnt);
  std::vector<std::pair<uint32_t, uint32_t>> tmp;  // (secondary key, suffix)
  for (size_t hh = h0; !groups.empty(); hh *= 2) {
    next_groups.clear();
    for (auto& g : groups) {
      size_t s = g.first, e = g.second;
      tmp.resize(e - s);
      for (size_t i = s; i < e; i++) {
        size_t p = (size_t)sa[i] + hh;
        tmp[i - s] = {p < n ? rnk[p] + 1 : 0u, sa[i]};
      }
      std::sort(tmp.begin(), tmp.end());
      for (size_t i = s; i < e; i++) sa[i] = tmp[i - s].second;
      size_t a = 0;
      while (a < tmp.size()) {
        size_t b = a + 1;
        while (b < tmp.size() && tmp[b].first == tmp[a].first) b++;
        for (size_t i = a; i < b; i++) rnk[tmp[i].second] = (uint32_t)(s + a);
        if (b - a > 1) next_groups.emplace_back((uint32_t)(s + a), (uint32_t)(s + b));
        a = b;
      }
    }
    groups.swap(next_groups);
  }
  return sa;
}

// ---------------------------------------------------------------------------------------------
// FM / FMD index pieces (rust-bio 0.37.1)  [recalled, unpinned]
// ---------------------------------------------------------------------------------------------
static inline int sym_code(uint8_t a) {
  switch (a) {
    case '$': return 0;
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 3;
    case 'N': return 4;
    case 'T': return 5;
    default: return -1;
  }
}

// bio Occ::get: checkpoint every occ_rate rows plus a byte count over the remainder
Counters& Index::tl() {
  static thread_local Counters c;
  return c;
}
void Index::fold_tl() const {
  Counters& c = tl();
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  counters.swg_cells += c.swg_cells; counters.swg_calls += c.swg_calls; counters.occ_lookups += c.occ_lookups;
  counters.fmd_ext += c.fmd_ext; counters.sa_locates += c.sa_locates; counters.hits += c.hits;
  c = Counters();
}

size_t Index::occ(size_t r, uint8_t a) const {
  tl().occ_lookups++;
  int c = sym_code(a);
  size_t i = r / occ_rate;
  size_t cnt = occ_samples[i * 6 + c];
  const uint8_t* p = bwt.data();
  for (size_t t = i * occ_rate + 1; t <= r; t++) cnt += (p[t] == a);
  return cnt;
}

// bio SampledSuffixArray::get: LF-walk to a sampled row (or a sentinel row kept as "extra")
size_t Index::sa_get(size_t r) const {
  tl().sa_locates++;
  size_t pos = r, offset = 0;
  for (;;) {
    if (pos % sa_rate == 0) return (size_t)sa_samples[pos / sa_rate] + offset;
    uint8_t c = bwt[pos];
    if (c == '$') return sa_extra.at(pos) + offset;
    pos = less[c] + occ(pos - 1, c);
    offset++;
  }
}

namespace {
struct BiInterval {
  size_t lower, lower_rev, size;
};
struct Fmd {
  const Index& ix;
  size_t less(uint8_t a) const { return ix.less[a]; }
  BiInterval init(uint8_t a) const {
    if (sym_code(a) < 1) return BiInterval{0, 0, 0};  // not ACGNT: matches nothing (out of domain)
    return BiInterval{less(a), less(complement(a)), ix.less[(size_t)a + 1] - ix.less[a]};
  }
  BiInterval backward_ext(const BiInterval& iv, uint8_t a) const {
    Index::tl().fmd_ext++;
    if (sym_code(a) < 0 || iv.size == 0) return BiInterval{0, 0, 0};
    size_t s = 0, o = 0, l = iv.lower_rev;
    static const uint8_t order[6] = {'$', 'T', 'G', 'C', 'N', 'A'};
    for (uint8_t b : order) {
      l += s;
      o = iv.lower == 0 ? 0 : ix.occ(iv.lower - 1, b);
      s = ix.occ(iv.lower + iv.size - 1, b) - o;
      if (b == a) break;
    }
    return BiInterval{less(a) + o, l, s};
  }
  static BiInterval swapped(const BiInterval& iv) { return BiInterval{iv.lower_rev, iv.lower, iv.size}; }
  BiInterval forward_ext(const BiInterval& iv, uint8_t a) const {
    return swapped(backward_ext(swapped(iv), complement(a)));
  }
  struct Smem { BiInterval iv; size_t pos, len; };
  // bio FMDIndex::smems(pattern, i, l)
  std::vector<Smem> smems(const uint8_t* P, size_t n, size_t i, size_t l) const {
    std::vector<std::pair<BiInterval, size_t>> curr, prev;
    std::vector<Smem> out;
    size_t match_len = 0;
    BiInterval interval = init(P[i]);
    if (interval.size != 0) match_len += 1;
    for (size_t t = i + 1; t < n; t++) {
      BiInterval f = forward_ext(interval, P[t]);
      if (interval.size != f.size) curr.push_back({interval, match_len});
      if (f.size == 0) break;
      interval = f;
      match_len += 1;
    }
    curr.push_back({interval, match_len});
    std::reverse(curr.begin(), curr.end());
    std::swap(curr, prev);
    long j = (long)n;
    for (long k = (long)i - 1; k >= -1; k--) {
      uint8_t a = (k == -1) ? (uint8_t)'$' : P[k];
      curr.clear();
      long last_size = -1;
      for (auto& pr : prev) {
        BiInterval f = backward_ext(pr.first, a);
        if ((f.size == 0 || k == -1) && curr.empty() && k < j && pr.second >= l) {
          j = k;
          out.push_back(Smem{pr.first, (size_t)(k + 1), pr.second});
        }
        if (f.size != 0 && (long)f.size != last_size) {
          last_size = (long)f.size;
          curr.push_back({f, pr.second + 1});
        }
      }
      if (curr.empty()) break;
      std::swap(curr, prev);
    }
    return out;
  }
  // bio FMDIndex::all_smems(pattern, l)
  std::vector<Smem> all_smems(const uint8_t* P, size_t n, size_t l) const {
    std::vector<Smem> out;
    size_t i0 = 0;
    while (i0 < n) {
      auto cur = smems(P, n, i0, l);
      size_t next = i0 + 1;
      for (auto& s : cur) next = std::max(next, s.pos + s.len);
      i0 = next;
      out.insert(out.end(), cur.begin(), cur.end());
    }
    return out;
  }
};
}  // namespace

// src/index.rs:228-255
std::vector<Mem> Index::all_smems(const uint8_t* query, size_t qlen, size_t min_seed_len) const {
  std::vector<Mem> mems;
  Fmd fmd{*this};
  auto intervals = fmd.all_smems(query, qlen, min_seed_len);
  for (auto& s : intervals) {
    if (s.iv.size == 0) continue;  // only reachable with min_seed_len == 0 on an unmatched symbol
    for (size_t r = s.iv.lower; r < s.iv.lower + s.iv.size; r++)  // interval.forward().occ(&sa)
      mems.push_back(Mem{sa_get(r), s.pos, s.len});
  }
  std::stable_sort(mems.begin(), mems.end(), [](const Mem& a, const Mem& b) { return a.len < b.len; });
  std::reverse(mems.begin(), mems.end());
  tl().hits += mems.size();
  return mems;
}

// Definition-based SMEMs (test cross-check for the recalled FMD algorithm): read intervals
// [q, E(q)) that occur in the text and are contained in no other occurring interval, every
// occurrence in suffix-array rank order, emitted in the order rule of SURVEY 8a-1.
std::vector<Mem> Index::all_smems_brute(const uint8_t* query, size_t qlen, size_t k) const {
  const size_t T = text.size();
  auto occurs_len = [&](size_t q, size_t p) {  // match length of query[q..] vs text[p..]
    size_t m = 0;
    while (q + m < qlen && p + m < T && text[p + m] == query[q + m] && text[p + m] != '$') m++;
    return m;
  };
  std::vector<size_t> E(qlen, 0);
  for (size_t q = 0; q < qlen; q++) {
    size_t best = 0;
    for (size_t p = 0; p < T; p++) best = std::max(best, occurs_len(q, p));
    E[q] = q + best;
  }
  struct S { size_t pos, len; };
  std::vector<S> smems;  // ascending start
  for (size_t q = 0; q < qlen; q++) {
    if (E[q] == q) continue;
    if (q > 0 && E[q - 1] >= E[q]) continue;
    if (E[q] - q >= k && E[q] - q >= 1) smems.push_back(S{q, E[q] - q});
  }
  // emission order: i0 = 0; emit smems covering i0 in descending start; i0 = max end or i0+1
  std::vector<S> emitted;
  size_t i0 = 0;
  while (i0 < qlen) {
    size_t next = i0 + 1;
    std::vector<S> cover;
    for (auto& s : smems) if (s.pos <= i0 && i0 < s.pos + s.len) cover.push_back(s);
    std::reverse(cover.begin(), cover.end());
    for (auto& s : cover) { emitted.push_back(s); next = std::max(next, s.pos + s.len); }
    i0 = next;
  }
  std::vector<Mem> mems;
  for (auto& s : emitted)
    for (size_t r = 0; r < T; r++) {
      size_t p = full_sa[r];
      if (occurs_len(s.pos, p) >= s.len) mems.push_back(Mem{p, s.pos, s.len});
    }
  std::stable_sort(mems.begin(), mems.end(), [](const Mem& a, const Mem& b) { return a.len < b.len; });
  std::reverse(mems.begin(), mems.end());
  return mems;
}

// src/index.rs:287-290
size_t Index::idx_to_ref(size_t idx) const {
  size_t lo = 0, hi = refs.size();
  while (lo < hi) {  // partition_point(|x| x.end_idx <= idx)
    size_t mid = (lo + hi) / 2;
    if (refs[mid].end_idx <= idx) lo = mid + 1; else hi = mid;
  }
  if (lo >= refs.size()) throw ReferencePanic("index.rs:289 index out of bounds");
  return lo;
}

// src/index.rs:304-323
std::vector<uint8_t> Index::seq_slice(size_t start, size_t end) const {
  size_t ri = idx_to_ref(start);
  const Ref& cur = refs[ri];
  if (cur.has_seq) {
    return std::vector<uint8_t>(cur.seq.begin() + (start - cur.start_idx),
                                cur.seq.begin() + (end - cur.start_idx));
  }
  const Ref& prev = refs[ri - 1];
  size_t cs = cur.end_idx - 1 - end, ce = cur.end_idx - 1 - start;
  return revcomp(prev.seq.data() + cs, ce - cs);
}

// ---------------------------------------------------------------------------------------------
// GTF -> transcriptome (cellranger `transcriptome` 0.1.0)  [recalled, unpinned]: genes and
// transcripts in file order, exons 0-based half-open sorted ascending, gene name falls back to id.
// ---------------------------------------------------------------------------------------------
namespace {
struct GtfTx {
  std::string id, chrom, gene_id;
  bool forward;
  std::vector<std::pair<size_t, size_t>> exons;
};
std::string gtf_attr(const std::string& attrs, const std::string& key) {
  size_t p = 0;
  while (p < attrs.size()) {
    while (p < attrs.size() && (attrs[p] == ' ' || attrs[p] == ';')) p++;
    size_t e = attrs.find(' ', p);
    if (e == std::string::npos) break;
    std::string k = attrs.substr(p, e - p);
    size_t v0 = e + 1, v1;
    std::string val;
    if (v0 < attrs.size() && attrs[v0] == '"') {
      v1 = attrs.find('"', v0 + 1);
      if (v1 == std::string::npos) break;
      val = attrs.substr(v0 + 1, v1 - v0 - 1);
      p = v1 + 1;
    } else {
      v1 = attrs.find(';', v0);
      if (v1 == std::string::npos) v1 = attrs.size();
      val = attrs.substr(v0, v1 - v0);
      p = v1;
    }
    if (k == key) return val;
  }
  return "";
}
}  // namespace

// src/index.rs:52-223
Index Index::create(const std::vector<FastaRecord>& fasta, const std::string& gtf_text,
                    size_t sa_sampling_rate, size_t occ_sampling_rate) {
  Index ix;
  ix.sa_rate = sa_sampling_rate;
  ix.occ_rate = occ_sampling_rate;
  std::vector<uint8_t>& seq = ix.text;
  std::map<std::pair<std::string, bool>, size_t> name_to_ref;
  std::unordered_map<std::string, size_t> chrom_fwd;
  for (auto& rec : fasta) {  // :67-101
    std::string name = rec.id.substr(0, rec.id.find(' '));
    size_t start_idx = seq.size();
    std::vector<uint8_t> cur(rec.seq);
    for (auto& c : cur) c = upper(c);
    seq.insert(seq.end(), cur.begin(), cur.end());
    seq.push_back('$');
    name_to_ref[{name, true}] = ix.refs.size();
    chrom_fwd[name] = ix.refs.size();
    ix.refs.push_back(Ref{name, true, cur, true, rec.seq.size(), start_idx, seq.size()});
    start_idx = seq.size();
    std::vector<uint8_t> rc = revcomp(rec.seq.data(), rec.seq.size());
    for (auto& c : rc) c = upper(c);
    seq.insert(seq.end(), rc.begin(), rc.end());
    seq.push_back('$');
    name_to_ref[{name, false}] = ix.refs.size();
    ix.refs.push_back(Ref{name, false, {}, false, rec.seq.size(), start_idx, seq.size()});
  }
  for (uint8_t c : seq)
    if (sym_code(c) < 0) throw std::runtime_error("reference contains a symbol outside ACGNT");

  // :103-111 SA, BWT, Less, Occ, sampled SA
  const size_t n = seq.size();
  ix.full_sa = suffix_array(seq);
  ix.bwt.resize(n);
  for (size_t r = 0; r < n; r++) ix.bwt[r] = ix.full_sa[r] > 0 ? seq[ix.full_sa[r] - 1] : (uint8_t)'$';
  ix.less.assign(257, 0);
  {
    size_t cnt[256] = {0};
    for (uint8_t c : ix.bwt) cnt[c]++;
    for (int c = 0; c < 256; c++) ix.less[c + 1] = ix.less[c] + cnt[c];
  }
  {
    uint32_t run[6] = {0, 0, 0, 0, 0, 0};
    ix.occ_samples.reserve((n / ix.occ_rate + 1) * 6);
    for (size_t r = 0; r < n; r++) {
      run[sym_code(ix.bwt[r])]++;
      if (r % ix.occ_rate == 0) ix.occ_samples.insert(ix.occ_samples.end(), run, run + 6);
    }
  }
  for (size_t r = 0; r < n; r++) {
    if (r % ix.sa_rate == 0) ix.sa_samples.push_back(ix.full_sa[r]);
    else if (ix.bwt[r] == '$') ix.sa_extra[r] = ix.full_sa[r];
  }

  // :115-124 transcriptome from GTF
  std::vector<GtfTx> gtxs;
  std::unordered_map<std::string, size_t> gene_by_id, tx_by_id;
  for (auto& ln : split_lines(gtf_text)) {
    if (ln.empty() || ln[0] == '#') continue;
    std::vector<std::string> f;
    size_t p = 0;
    for (int t = 0; t < 8; t++) {
      size_t e = ln.find('\t', p);
      if (e == std::string::npos) { p = std::string::npos; break; }
      f.push_back(ln.substr(p, e - p));
      p = e + 1;
    }
    if (p == std::string::npos) continue;
    f.push_back(ln.substr(p));
    const std::string& feat = f[2];
    std::string gene_id = gtf_attr(f[8], "gene_id");
    auto ensure_gene = [&]() {
      auto it = gene_by_id.find(gene_id);
      if (it != gene_by_id.end()) return it->second;
      std::string gname = gtf_attr(f[8], "gene_name");
      gene_by_id[gene_id] = ix.txome.genes.size();
      ix.txome.genes.push_back(Gene{gene_id, gname.empty() ? gene_id : gname});
      return ix.txome.genes.size() - 1;
    };
    auto ensure_tx = [&]() {
      std::string tid = gtf_attr(f[8], "transcript_id");
      auto it = tx_by_id.find(tid);
      if (it != tx_by_id.end()) return it->second;
      ensure_gene();
      tx_by_id[tid] = gtxs.size();
      gtxs.push_back(GtfTx{tid, f[0], gene_id, f[6] != "-", {}});
      return gtxs.size() - 1;
    };
    if (feat == "gene") ensure_gene();
    else if (feat == "transcript") ensure_tx();
    else if (feat == "exon") {
      size_t t = ensure_tx();
      gtxs[t].exons.push_back({(size_t)std::stoull(f[3]) - 1, (size_t)std::stoull(f[4])});
    }
  }
  for (auto& t : gtxs) std::sort(t.exons.begin(), t.exons.end());

  // :126-206
  std::vector<std::pair<size_t, size_t>> gene_iv(ix.txome.genes.size(), {n, 0});  // :134
  for (size_t ti = 0; ti < gtxs.size(); ti++) {
    GtfTx& tx = gtxs[ti];
    if (tx.exons.empty()) throw std::runtime_error("transcript without exons: " + tx.id);
    size_t gene_idx = gene_by_id.at(tx.gene_id);
    const Ref& fwd = ix.refs[chrom_fwd.at(tx.chrom)];
    std::vector<uint8_t> tx_seq;  // tx.get_sequence: exon slices, reverse-complemented for '-'
    for (auto& e : tx.exons) {
      if (e.second > fwd.seq.size() || e.first >= e.second) throw std::runtime_error("exon outside its sequence: " + tx.id);
      tx_seq.insert(tx_seq.end(), fwd.seq.begin() + e.first, fwd.seq.begin() + e.second);
    }
    if (!tx.forward) tx_seq = revcomp(tx_seq.data(), tx_seq.size());
    for (auto& c : tx_seq) c = upper(c);
    bool strand = tx.forward;
    const Ref& tx_ref = ix.refs[name_to_ref.at({tx.chrom, strand})];
    size_t t_start = tx.exons.front().first, t_end = tx.exons.back().second;
    size_t tx_start = strand ? t_start + tx_ref.start_idx : tx_ref.end_idx - 1 - t_end;
    size_t tx_end = strand ? t_end + tx_ref.start_idx : tx_ref.end_idx - 1 - t_start;
    gene_iv[gene_idx] = {std::min(gene_iv[gene_idx].first, tx_start), std::max(gene_iv[gene_idx].second, tx_end)};
    std::vector<Exon> exons;
    for (auto& e : tx.exons) {  // :164-191
      size_t es = strand ? e.first + tx_ref.start_idx : tx_ref.end_idx - 1 - e.second;
      size_t ee = strand ? e.second + tx_ref.start_idx : tx_ref.end_idx - 1 - e.first;
      ix.txome.exon_to_tx.insert(es, ee, ti);
      exons.push_back(Exon{es, ee, ti});
    }
    if (!strand) std::reverse(exons.begin(), exons.end());  // :192-195
    ix.txome.txs.push_back(Tx{tx.id, tx.chrom, strand, exons, tx_seq, gene_idx});
  }
  for (size_t g = 0; g < gene_iv.size(); g++)  // :208-213
    ix.txome.gene_intervals.insert(gene_iv[g].first, gene_iv[g].second, g);
  return ix;
}

Index Index::create_from_files(const std::string& ref_path, const std::string& annot_path,
                               size_t sa_rate, size_t occ_rate) {
  return create(parse_fasta(read_file(ref_path)), read_file(annot_path), sa_rate, occ_rate);
}

// ---------------------------------------------------------------------------------------------
// txome lifting (src/txome.rs)
// ---------------------------------------------------------------------------------------------
// src/txome.rs:77-79
bool intersect(size_t a0, size_t a1, size_t b0, size_t b1) {
  return (a0 >= b0 && a0 < b1) || (b0 >= a0 && b0 < a1);
}

// src/txome.rs:82-103
Mem lift_mem_to_tx(const Mem& mem, const Tx& tx) {
  size_t exon_sum = 0;
  for (auto& exon : tx.exons) {
    if (intersect(mem.ref_idx, mem.ref_idx + mem.len, exon.start, exon.end)) {
      size_t start = sat_sub(mem.ref_idx, exon.start) + exon_sum;
      size_t start_offset = sat_sub(exon.start, mem.ref_idx);
      size_t end = std::min(mem.ref_idx + mem.len, exon.end) - exon.start + exon_sum;
      return Mem{start, mem.query_idx + start_offset, end - start};
    }
    exon_sum += exon.len();
  }
  throw ReferencePanic("txome.rs:102 unreachable");
}

// src/txome.rs:110-160
Alignment lift_tx_to_gx(const Alignment& tx_aln, const Tx& tx) {
  Alignment aln = tx_aln;
  aln.operations.clear();
  size_t i = tx_aln.ystart, op_idx = 0, exon_sum = 0, exon_idx = 0;
  while (exon_sum + tx.exons.at(exon_idx).len() <= i) {
    exon_sum += tx.exons[exon_idx].len();
    exon_idx++;
  }
  aln.ystart = tx.exons[exon_idx].start + (i - exon_sum);
  while (op_idx < tx_aln.operations.size()) {
    if (exon_idx + 1 < tx.exons.size() && exon_sum + tx.exons[exon_idx].len() <= i) {
      exon_sum += tx.exons[exon_idx].len();
      exon_idx++;
      aln.operations.push_back(Op{Yclip, tx.exons[exon_idx].start - tx.exons[exon_idx - 1].end});
    }
    uint8_t k = tx_aln.operations[op_idx].kind;
    if (k == Match || k == Subst || k == Del) i++;
    aln.operations.push_back(tx_aln.operations[op_idx]);
    op_idx++;
  }
  if (i != tx_aln.yend) throw ReferencePanic("txome.rs:154 assert_eq!(i, tx_aln.yend)");
  aln.yend = tx.exons[exon_idx].start + (i - exon_sum);
  return aln;
}

// ---------------------------------------------------------------------------------------------
// aligner (src/aligner.rs)
// ---------------------------------------------------------------------------------------------
// src/aligner.rs:410-426
void extend_seed_match(const uint8_t* ref_seq, size_t ref_len, Mem& hit, const uint8_t* read, size_t read_len) {
  while (hit.ref_idx + hit.len < ref_len && hit.query_idx + hit.len < read_len &&
         ref_seq[hit.ref_idx + hit.len] == read[hit.query_idx + hit.len])
    hit.len++;
  while (hit.ref_idx > 0 && hit.query_idx > 0 && ref_seq[hit.ref_idx - 1] == read[hit.query_idx - 1]) {
    hit.ref_idx--;
    hit.query_idx--;
    hit.len++;
  }
}

// src/aligner.rs:352-407
Alignment extend_left_right(const uint8_t* ref_seq, size_t ref_len, const Mem& hit, const uint8_t* read,
                            size_t read_len, SwgExtend& swg, size_t band_width, int32_t x_drop) {
  size_t xo = hit.query_idx + hit.len, yo = hit.ref_idx + hit.len;
  Alignment right = swg.extend(read + xo, read_len - xo, ref_seq + yo, ref_len - yo, band_width, x_drop);
  std::vector<uint8_t> x(read, read + hit.query_idx);
  std::reverse(x.begin(), x.end());
  size_t ys = sat_sub(hit.ref_idx, read_len + band_width);
  std::vector<uint8_t> y(ref_seq + ys, ref_seq + hit.ref_idx);
  std::reverse(y.begin(), y.end());
  Alignment left = swg.extend(x.data(), x.size(), y.data(), y.size(), band_width, x_drop);

  Alignment a;
  a.ystart = hit.ref_idx - left.yend;
  a.yend = hit.ref_idx + hit.len + right.yend;
  a.xstart = hit.query_idx - left.xend;
  a.xend = hit.query_idx + hit.len + right.xend;
  a.score = left.score + (int32_t)hit.len + right.score;
  a.operations.assign(left.operations.rbegin(), left.operations.rend());
  for (size_t t = 0; t < hit.len; t++) a.operations.push_back(Op{Match, 1});
  a.operations.insert(a.operations.end(), right.operations.begin(), right.operations.end());
  a.ylen = ref_len;
  a.xlen = read_len;
  return a;
}

// src/aligner.rs:429-449
static Alignment concat_to_chr_aln(const Index& index, Alignment aln) {
  const Ref& r = index.refs[index.idx_to_ref(aln.ystart)];
  if (r.strand) {
    aln.ystart -= r.start_idx;
    aln.yend -= r.start_idx;
    aln.ylen = r.len;
  } else {
    size_t ys = r.len - (aln.yend - r.start_idx), ye = r.len - (aln.ystart - r.start_idx);
    aln.ystart = ys;
    aln.yend = ye;
    aln.ylen = r.len;
    std::reverse(aln.operations.begin(), aln.operations.end());
  }
  return aln;
}

// src/aligner.rs:198-314
GenomeAlignment align_seed_hit(const Index& index, const std::vector<uint8_t>& read, const Mem& hit,
                               SwgExtend& swg, size_t band_width, int32_t x_drop) {
  size_t ref_id = index.idx_to_ref(hit.ref_idx);
  const Ref& aln_ref = index.refs[ref_id];
  Alignment gx_aln;
  {
    size_t seq_start = std::max(sat_sub(hit.ref_idx, read.size() + band_width), aln_ref.start_idx);
    size_t seq_end = std::min(hit.ref_idx + hit.len + read.size() + band_width, aln_ref.end_idx - 1);
    std::vector<uint8_t> ref_seq = index.seq_slice(seq_start, seq_end);
    Mem rel = hit;
    rel.ref_idx -= seq_start;
    gx_aln = extend_left_right(ref_seq.data(), ref_seq.size(), rel, read.data(), read.size(), swg, band_width, x_drop);
    gx_aln.ystart += seq_start;
    gx_aln.yend += seq_start;
  }
  bool have_tx = false;
  size_t best_tx = 0;
  Alignment best_tx_aln;
  for (size_t tx_idx : index.txome.exon_to_tx.find(hit.ref_idx, hit.ref_idx + hit.len)) {
    const Tx& tx = index.txome.txs[tx_idx];
    Mem tx_seed = lift_mem_to_tx(hit, tx);
    extend_seed_match(tx.seq.data(), tx.seq.size(), tx_seed, read.data(), read.size());
    Alignment tx_aln = extend_left_right(tx.seq.data(), tx.seq.size(), tx_seed, read.data(), read.size(), swg, band_width, x_drop);
    int32_t s = tx_aln.score;
    if (!have_tx || s > best_tx_aln.score) {
      have_tx = true;
      best_tx = tx_idx;
      best_tx_aln = std::move(tx_aln);
    }
    if (s >= (int32_t)read.size()) break;  // :253-257
  }
  GenomeAlignment g;
  g.ref_name = aln_ref.name;
  g.ref_id = ref_id;
  g.strand = aln_ref.strand;
  g.primary = false;
  if (have_tx && best_tx_aln.score >= gx_aln.score) {  // :263 ties -> Exonic
    Alignment lifted = lift_tx_to_gx(best_tx_aln, index.txome.txs[best_tx]);
    g.gx_aln = concat_to_chr_aln(index, std::move(lifted));
    g.aln_type = Exonic;
    g.tx_aln = std::move(best_tx_aln);
    g.tx_idx = best_tx;
  } else {
    std::vector<size_t> genes = index.txome.gene_intervals.find(gx_aln.ystart, gx_aln.yend);
    g.gx_aln = concat_to_chr_aln(index, std::move(gx_aln));
    if (genes.empty()) g.aln_type = Intergenic;
    else { g.aln_type = Intronic; g.gene_idx = genes[0]; }
  }
  return g;
}

// src/aligner.rs:317-349
std::vector<GenomeAlignment> filter_overlapping(std::vector<GenomeAlignment> alns) {
  if (alns.empty()) return alns;
  std::stable_sort(alns.begin(), alns.end(), [](const GenomeAlignment& a, const GenomeAlignment& b) {
    if (a.ref_name != b.ref_name) return a.ref_name < b.ref_name;
    if (a.strand != b.strand) return (int)a.strand < (int)b.strand;
    return a.gx_aln.ystart < b.gx_aln.ystart;
  });
  size_t max_end = 0;
  std::vector<GenomeAlignment> res;
  for (auto& aln : alns) {
    if (aln.gx_aln.ystart >= max_end || aln.ref_name != res.back().ref_name || aln.strand != res.back().strand) {
      max_end = aln.gx_aln.yend;
      res.push_back(std::move(aln));
    } else {
      GenomeAlignment& cur = res.back();
      if (aln.gx_aln.score > cur.gx_aln.score) cur = std::move(aln);
      max_end = std::max(max_end, cur.gx_aln.yend);
    }
  }
  return res;
}

// src/aligner.rs:123-190
std::vector<GenomeAlignment> align_read(const Index& index, const uint8_t* read_in, size_t len, const AlignOpts& opts) {
  std::vector<uint8_t> read(read_in, read_in + len);
  for (auto& c : read) c = upper(c);  // :125
  std::vector<Mem> mems = index.all_smems(read.data(), read.size(), opts.min_seed_len);
  std::vector<GenomeAlignment> gx_alns;
  float prod = opts.min_aln_score_percent * (float)read.size();
  int32_t min_aln_score = std::max((int32_t)prod, opts.min_aln_score);  // :130-133
  int32_t max_aln_score = min_aln_score;
  size_t band_width = sat_sub(read.size(), (size_t)(int64_t)min_aln_score);  // :137 (`as usize` wraps)
  size_t x_drop = band_width;
  const int32_t range = (int32_t)opts.multimap_score_range;
  SwgExtend swg(band_width, -1, -1, 1, -1);  // :140-141
  uint64_t cells0 = swg.cells;
  for (auto& hit : mems) {
    GenomeAlignment g = align_seed_hit(index, read, hit, swg, band_width, (int32_t)x_drop);
    if (!opts.intron_mode && g.aln_type != Exonic) continue;  // :146-151
    int32_t s = g.gx_aln.score;
    if (s < opts.min_aln_score || s < min_aln_score || s < max_aln_score - range) continue;  // :154-159
    size_t lim = sat_sub(read.size() + opts.multimap_score_range, (size_t)(int64_t)s);  // :162-171
    band_width = std::min(band_width, lim);
    x_drop = std::min(x_drop, lim);
    max_aln_score = std::max(max_aln_score, s);
    gx_alns.push_back(std::move(g));
  }
  Index::tl().swg_cells += swg.cells - cells0;
  {  // :177-179
    std::vector<GenomeAlignment> kept;
    for (auto& g : gx_alns) if (g.gx_aln.score >= max_aln_score - range) kept.push_back(std::move(g));
    gx_alns.swap(kept);
  }
  gx_alns = filter_overlapping(std::move(gx_alns));  // :181
  std::stable_sort(gx_alns.begin(), gx_alns.end(), [](const GenomeAlignment& a, const GenomeAlignment& b) {
    return -a.gx_aln.score < -b.gx_aln.score;
  });  // :183
  if (!gx_alns.empty()) gx_alns[0].primary = true;  // :185-187
  return gx_alns;
}

// ---------------------------------------------------------------------------------------------
// writers (src/aln_writer.rs)
// ---------------------------------------------------------------------------------------------
// src/aln_writer.rs:332-340
uint8_t multimapq(size_t n) {
  if (n <= 1) return 255;
  if (n >= 5) return 0;
  float v = -10.0f * std::log10(1.0f - 1.0f / (float)n);
  return (uint8_t)std::lround(v);
}

// src/aln_writer.rs:279-323
std::string cigar_string(const std::vector<Op>& ops) {
  std::string out;
  int prev = -1;
  uint64_t prev_n = 0, prev_len = 0;
  auto flush = [&]() {
    if (prev < 0) return;
    static const char sym[6] = {'M', 'M', 'D', 'I', 'S', 'N'};
    uint64_t l = (prev == Xclip || prev == Yclip) ? prev_n : prev_len;
    out += std::to_string(l);
    out += sym[prev];
  };
  for (auto op : ops) {
    uint8_t k = op.kind == Subst ? (uint8_t)Match : op.kind;
    bool same = prev >= 0 && prev == (int)k && ((k != Xclip && k != Yclip) || prev_n == op.n);
    if (!same) {
      flush();
      prev = k;
      prev_n = op.n;
      prev_len = 1;
    } else {
      prev_len++;  // NB: equal adjacent clips would print their own length once (reference :292-294)
    }
  }
  flush();
  return out;
}

// src/aln_writer.rs:47-116
std::string paf_line(const std::string& qname, size_t qlen, const GenomeAlignment& aln, size_t multimap) {
  size_t num_match = 0, num_match_gap = 0;
  for (auto& op : aln.gx_aln.operations) {
    if (op.kind == Match) num_match++;
    if (op.kind != Yclip) num_match_gap++;
  }
  std::ostringstream s;
  s << qname << '\t' << qlen << '\t' << aln.gx_aln.xstart << '\t' << aln.gx_aln.xend << '\t'
    << (aln.strand ? "+" : "-") << '\t' << aln.ref_name << '\t' << aln.gx_aln.ylen << '\t'
    << aln.gx_aln.ystart << '\t' << aln.gx_aln.yend << '\t' << num_match << '\t' << num_match_gap << '\t'
    << (unsigned)multimapq(multimap) << '\t' << '\n';
  return s.str();
}

static std::string read_name(const std::string& id) { return id.substr(0, id.find(' ')); }  // :344-349
static std::string maybe_empty(const std::vector<uint8_t>& s) {                              // :352-358
  return s.empty() ? std::string("*") : std::string(s.begin(), s.end());
}

// src/aln_writer.rs:256-276 (noodles 0.1.0 header text: unpinned)
std::string sam_header(const Index& index) {
  std::ostringstream s;
  std::vector<std::string> seen;
  for (auto& r : index.refs) {
    if (std::find(seen.begin(), seen.end(), r.name) != seen.end()) continue;  // map keyed by name
    seen.push_back(r.name);
    s << "@SQ\tSN:" << r.name << "\tLN:" << r.len << '\n';
  }
  s << "@PG\tID:thermite\n";
  return s.str();
}

// src/aln_writer.rs:118-238 (noodles 0.1.0 record text: unpinned)
std::string sam_line(const Index& index, const FastqRecord& rec, const GenomeAlignment& aln, size_t multimap, size_t hit_index) {
  std::vector<uint8_t> seq = aln.strand ? rec.seq : revcomp(rec.seq.data(), rec.seq.size());
  std::vector<uint8_t> qual = rec.qual;
  if (!aln.strand) std::reverse(qual.begin(), qual.end());
  int flags = 0;
  if (!aln.strand) flags |= 0x10;
  if (!aln.primary) flags |= 0x100;
  size_t nm = 0;
  for (auto& op : aln.gx_aln.operations) if (op.kind == Subst) nm++;
  std::ostringstream s;
  s << read_name(rec.id) << '\t' << flags << '\t' << aln.ref_name << '\t' << (aln.gx_aln.ystart + 1) << '\t'
    << (unsigned)multimapq(multimap) << '\t' << cigar_string(aln.gx_aln.operations) << "\t*\t0\t0\t"
    << maybe_empty(seq) << '\t' << maybe_empty(qual) << "\tAS:i:" << aln.gx_aln.score << "\tNH:i:" << multimap
    << "\tHI:i:" << hit_index << "\tnM:i:" << nm;
  if (aln.aln_type == Exonic) {
    const Tx& tx = index.txome.txs[aln.tx_idx];
    const Gene& g = index.txome.genes[tx.gene_idx];
    s << "\tTX:Z:" << tx.id << ",+" << aln.tx_aln.ystart << ',' << cigar_string(aln.tx_aln.operations)
      << "\tGX:Z:" << g.id << "\tGN:Z:" << g.name << "\tRE:A:E";
  } else if (aln.aln_type == Intronic) {
    const Gene& g = index.txome.genes[aln.gene_idx];
    s << "\tGX:Z:" << g.id << "\tGN:Z:" << g.name << "\tRE:A:N";
  } else {
    s << "\tRE:A:I";
  }
  s << '\n';
  return s.str();
}

// src/aln_writer.rs:241-253
std::string sam_unmapped_line(const FastqRecord& rec) {
  std::ostringstream s;
  s << read_name(rec.id) << "\t4\t*\t0\t255\t*\t*\t0\t0\t" << maybe_empty(rec.seq) << '\t' << maybe_empty(rec.qual) << '\n';
  return s.str();
}

// src/aligner.rs:22-120
std::string align_fastq(const Index& index, const std::vector<FastqRecord>& reads, const AlignOpts& opts, bool sam) {
  std::string out;
  if (sam) out += sam_header(index);
  for (auto& rec : reads) {
    auto alns = align_read(index, rec.seq.data(), rec.seq.size(), opts);
    if (alns.empty()) {
      if (sam) out += sam_unmapped_line(rec);
      continue;
    }
    for (size_t i = 0; i < alns.size(); i++) {
      if (sam) out += sam_line(index, rec, alns[i], alns.size(), i + 1);
      else out += paf_line(rec.id, rec.seq.size(), alns[i], alns.size());
    }
  }
  return out;
}

}  // namespace orc
