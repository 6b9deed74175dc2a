// zq_dist.cpp -- the multi-GPU side of the C ABI (include/zq_b200.h, zq_dist_*): one process per GPU, the units
// (compressBlock calls, files) are dealt to the ranks and never move; what crosses NVLink is only
//   * the compressed size of every block (4 B/unit) -> every rank knows the archive offset of every block
//     (the counterpart of the writer thread's running offset, Z:71445-71518), and
//   * the 20-byte SHA-1 of every fragment -> every rank knows which of its fragments are new to the archive
//     (the job of the HTIndex lookup in Jidac::add, Z:122569-122573, Z:71567-71604).
// Transport: NCCL (ncclAllGather on the rank's device, library found with dlopen so that nothing links against it
// when one GPU is used) -- or a caller-supplied all-gather callback, which is how the world_size-2 tests run this
// host logic over gloo on CPUs.  No collective sits inside a unit; there is nothing to overlap with.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/zq_b200.h"

namespace {

struct NcclId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_getuid)(NcclId*);
typedef int (*fn_initrank)(NcclComm*, int, NcclId, int);
typedef int (*fn_allgather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_destroy)(NcclComm);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* h = nullptr;
  fn_getuid getuid = nullptr; fn_initrank initrank = nullptr; fn_allgather allgather = nullptr; fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
  std::string why;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { why = "libnccl.so.2 not found (dlopen)"; return false; }
    getuid = (fn_getuid)dlsym(h, "ncclGetUniqueId"); initrank = (fn_initrank)dlsym(h, "ncclCommInitRank");
    allgather = (fn_allgather)dlsym(h, "ncclAllGather"); destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (!getuid || !initrank || !allgather || !destroy) { why = "NCCL symbols missing"; h = nullptr; return false; }
    return true;
  }
};
NcclApi& nccl() { static NcclApi a; return a; }

thread_local std::string g_dist_error;

}  // namespace

struct zq_dist {
  int rank = 0, world = 1, device = -1;
  NcclComm comm = nullptr;
  cudaStream_t stream = nullptr;
  zq_allgather_fn cb = nullptr; void* cb_user = nullptr;
  void* d_send = nullptr; void* d_recv = nullptr; size_t send_cap = 0, recv_cap = 0;
  std::string err;
  uint64_t bytes_exchanged = 0;

  int fail(int code, const std::string& m) { err = m; return code; }

  // every rank contributes `bytes` bytes; out = world * bytes, rank order
  int allgather(const void* in, void* out, size_t bytes) {
    bytes_exchanged += bytes * (size_t)world;
    if (world == 1) { memcpy(out, in, bytes); return ZQ_OK; }
    if (cb) return cb(cb_user, in, out, bytes) == 0 ? ZQ_OK : fail(ZQ_E_NODEVICE, "all-gather callback failed");
    if (!comm) return fail(ZQ_E_NODEVICE, "no transport");
    cudaSetDevice(device);
    if (bytes > send_cap) { if (d_send) cudaFree(d_send); if (cudaMalloc(&d_send, bytes + 256) != cudaSuccess) return fail(ZQ_E_NOMEM, "Out of memory"); send_cap = bytes + 256; }
    if (bytes * world > recv_cap) { if (d_recv) cudaFree(d_recv); if (cudaMalloc(&d_recv, bytes * world + 256) != cudaSuccess) return fail(ZQ_E_NOMEM, "Out of memory"); recv_cap = bytes * world + 256; }
    if (bytes) cudaMemcpyAsync(d_send, in, bytes, cudaMemcpyHostToDevice, stream);
    const int rc = nccl().allgather(d_send, d_recv, bytes, /*ncclUint8*/ 1, comm, stream);
    if (rc != 0) return fail(ZQ_E_NODEVICE, std::string("ncclAllGather: ") + (nccl().errstr ? nccl().errstr(rc) : "error"));
    if (bytes) cudaMemcpyAsync(out, d_recv, bytes * world, cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return fail(ZQ_E_NODEVICE, "CUDA error in all-gather");
    return ZQ_OK;
  }
  // variable-length form: counts first, then payloads padded to the largest
  int allgatherv(const void* in, size_t bytes, std::vector<uint8_t>& out, std::vector<uint64_t>& sizes) {
    sizes.assign(world, 0);
    const uint64_t mine = bytes;
    int rc = allgather(&mine, sizes.data(), 8);
    if (rc) return rc;
    const size_t mx = (size_t)*std::max_element(sizes.begin(), sizes.end());
    std::vector<uint8_t> pad(mx, 0), all(mx * world);
    if (bytes) memcpy(pad.data(), in, bytes);
    if (mx) { rc = allgather(pad.data(), all.data(), mx); if (rc) return rc; }
    size_t tot = 0;
    for (uint64_t s : sizes) tot += (size_t)s;
    out.resize(tot);
    size_t at = 0;
    for (int r = 0; r < world; ++r) { if (sizes[r]) memcpy(out.data() + at, all.data() + (size_t)r * mx, (size_t)sizes[r]); at += (size_t)sizes[r]; }
    return ZQ_OK;
  }
};

extern "C" {

int zq_dist_unique_id(uint8_t out[128]) {
  if (!nccl().load()) { g_dist_error = nccl().why; return ZQ_E_NODEVICE; }
  NcclId id;
  const int rc = nccl().getuid(&id);
  if (rc != 0) { g_dist_error = "ncclGetUniqueId failed"; return ZQ_E_NODEVICE; }
  memcpy(out, id.internal, 128);
  return ZQ_OK;
}

zq_dist* zq_dist_create(int device, int rank, int world, const uint8_t id[128]) {
  if (world < 1 || rank < 0 || rank >= world) { g_dist_error = "bad rank / world"; return nullptr; }
  zq_dist* d = new zq_dist;
  d->rank = rank; d->world = world; d->device = device;
  if (world == 1) return d;
  if (!id) { g_dist_error = "world > 1 needs the unique id of rank 0 (zq_dist_unique_id)"; delete d; return nullptr; }
  if (!nccl().load()) { g_dist_error = nccl().why; delete d; return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_dist_error = "no CUDA device for the communicator"; delete d; return nullptr;
  }
  NcclId nid; memcpy(nid.internal, id, 128);
  const int rc = nccl().initrank(&d->comm, world, nid, rank);
  if (rc != 0) { g_dist_error = std::string("ncclCommInitRank: ") + (nccl().errstr ? nccl().errstr(rc) : "error"); cudaStreamDestroy(d->stream); delete d; return nullptr; }
  return d;
}

zq_dist* zq_dist_create_cb(int rank, int world, zq_allgather_fn fn, void* user) {
  if (world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) { g_dist_error = "bad rank / world / callback"; return nullptr; }
  zq_dist* d = new zq_dist;
  d->rank = rank; d->world = world; d->cb = fn; d->cb_user = user;
  return d;
}

void zq_dist_destroy(zq_dist* d) {
  if (!d) return;
  if (d->comm) nccl().destroy(d->comm);
  if (d->d_send) cudaFree(d->d_send);
  if (d->d_recv) cudaFree(d->d_recv);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

const char* zq_dist_last_error(zq_dist* d) { return d ? d->err.c_str() : g_dist_error.c_str(); }
uint64_t zq_dist_bytes_exchanged(zq_dist* d) { return d ? d->bytes_exchanged : 0; }

// contiguous, balanced [lo, hi) of `total` units for `rank` (sizes differ by at most one)
int zq_dist_shard_range(uint64_t total, int rank, int world, uint64_t* lo, uint64_t* hi) {
  if (world < 1 || rank < 0 || rank >= world || !lo || !hi) return ZQ_E_ARG;
  const uint64_t base = total / world, rem = total % world;
  *lo = rank * base + std::min<uint64_t>(rank, rem);
  *hi = *lo + base + ((uint64_t)rank < rem ? 1 : 0);
  return ZQ_OK;
}

// longest-processing-time assignment of units with unequal predicted cost (bytes x method weight): deterministic,
// every rank computes the same owner[]
int zq_dist_shard_lpt(const uint64_t* cost, uint64_t n, int world, int32_t* owner) {
  if (world < 1 || (n && (!cost || !owner))) return ZQ_E_ARG;
  std::vector<uint64_t> order(n);
  for (uint64_t i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return cost[a] > cost[b]; });
  std::vector<uint64_t> load(world, 0);
  for (uint64_t i : order) {
    int best = 0;
    for (int r = 1; r < world; ++r) if (load[r] < load[best]) best = r;
    owner[i] = best; load[best] += cost[i];
  }
  return ZQ_OK;
}

// The ranks hold the compressed sizes of their contiguous shards (zq_dist_shard_range of `total`): afterwards every
// rank has all sizes and the byte offset of every block in the archive stream (exclusive scan in unit order).
int zq_dist_exchange_sizes(zq_dist* d, const uint32_t* local_sizes, uint64_t total, uint32_t* all_sizes, uint64_t* offsets) {
  if (!d) return ZQ_E_NODEVICE;
  if (total && (!all_sizes || !offsets)) return d->fail(ZQ_E_ARG, "bad argument");
  uint64_t lo, hi;
  zq_dist_shard_range(total, d->rank, d->world, &lo, &hi);
  if (hi > lo && !local_sizes) return d->fail(ZQ_E_ARG, "bad argument");
  const uint64_t width = (total + d->world - 1) / d->world;      // every rank sends the same number of slots
  std::vector<uint32_t> send(std::max<uint64_t>(width, 1), 0), recv(std::max<uint64_t>(width, 1) * d->world, 0);
  for (uint64_t k = 0; k < hi - lo; ++k) send[k] = local_sizes[k];
  if (width) { const int rc = d->allgather(send.data(), recv.data(), (size_t)width * 4); if (rc) return rc; }
  uint64_t at = 0;
  for (int r = 0; r < d->world; ++r) {
    uint64_t a, b;
    zq_dist_shard_range(total, r, d->world, &a, &b);
    for (uint64_t k = 0; k < b - a; ++k) all_sizes[a + k] = recv[(uint64_t)r * width + k];
  }
  for (uint64_t u = 0; u < total; ++u) { offsets[u] = at; at += all_sizes[u]; }
  return ZQ_OK;
}

// Dedup keys: every rank contributes the 20-byte digests of its fragments, in archive order within the rank.  A
// fragment is "first" if no rank before this one, and no earlier fragment of this rank, has the same digest -- the
// archive stores it; every other occurrence becomes a reference.  unique_total = fragments stored by all ranks.
int zq_dist_dedup(zq_dist* d, const uint8_t* local_digests, uint64_t n_local, uint8_t* is_first, uint64_t* unique_total) {
  if (!d) return ZQ_E_NODEVICE;
  if (n_local && (!local_digests || !is_first)) return d->fail(ZQ_E_ARG, "bad argument");
  std::vector<uint8_t> all; std::vector<uint64_t> sizes;
  const int rc = d->allgatherv(local_digests, (size_t)n_local * 20, all, sizes);
  if (rc) return rc;
  struct Key { uint64_t a, b; uint32_t c; bool operator==(const Key& o) const { return a == o.a && b == o.b && c == o.c; } };
  struct KeyHash { size_t operator()(const Key& k) const { return (size_t)(k.a ^ (k.b * 0x9E3779B97F4A7C15ull) ^ k.c); } };
  std::unordered_map<Key, uint64_t, KeyHash> seen;   // digest -> global index of its first occurrence
  seen.reserve(all.size() / 20 + 16);
  uint64_t g = 0, my_first = 0;
  for (int r = 0; r < d->rank; ++r) my_first += sizes[r] / 20;
  for (size_t at = 0; at + 20 <= all.size(); at += 20, ++g) {
    Key k; memcpy(&k.a, &all[at], 8); memcpy(&k.b, &all[at + 8], 8); memcpy(&k.c, &all[at + 16], 4);
    const bool fresh = seen.emplace(k, g).second;
    if (g >= my_first && g < my_first + n_local) is_first[g - my_first] = fresh ? 1 : 0;
  }
  if (unique_total) *unique_total = seen.size();
  return ZQ_OK;
}
// One stream cut across the ranks (a file larger than one GPU's share).  Rank r owns bytes
// [lo, hi) and also holds the bytes up to avail_end (its right neighbour's first ones).  It has fragmented
// [from, avail_end) with a fresh chunker started at `from` (zq_fragment on that piece): a SPECULATIVE chain unless
// `from` is a boundary of the stream's real chain.  The chunker forgets everything at a fragment boundary
// (Z:122457-122470), so where the real chain, arriving from the left neighbour, ends a fragment at an offset at which
// this rank's chain also ends one (or starts), the two are identical from there on.  Every rank computes the same
// decisions from the gathered boundary lists, left to right:
//   s_0 = 0;  s_r = the first real boundary >= lo_r of rank r-1's chain that rank r's chain shares.
// Rank r keeps its fragments starting in [s_r, s_{r+1}): the fragment straddling hi_r belongs to the left owner.
// If the chains do not meet inside the overlap (a constant run -- zero pages -- never re-synchronises), the rank is
// told the first real boundary >= lo_r (restart_at): it fragments again from there and every rank calls once more.
int zq_dist_stitch_fragments(zq_dist* d, uint64_t stream_total, uint64_t lo, uint64_t hi, uint64_t from, uint64_t avail_end,
                             const uint32_t* frag_len, uint64_t nfrag, zq_stitch* out) {
  if (!d) return ZQ_E_NODEVICE;
  if (!out || (nfrag && !frag_len) || lo > hi || hi > avail_end || avail_end > stream_total || from < lo || from > avail_end)
    return d->fail(ZQ_E_ARG, "bad argument");
  memset(out, 0, sizeof(*out));
  const int W = d->world;
  // header: lo, hi, from, avail_end, nfrag; then the end offset of every fragment
  std::vector<uint64_t> mine(5 + nfrag);
  mine[0] = lo; mine[1] = hi; mine[2] = from; mine[3] = avail_end; mine[4] = nfrag;
  uint64_t at = from;
  for (uint64_t k = 0; k < nfrag; ++k) { at += frag_len[k]; mine[5 + k] = at; }
  if (at != avail_end) return d->fail(ZQ_E_ARG, "fragment lengths do not add up to the piece");
  std::vector<uint8_t> all; std::vector<uint64_t> sizes;
  const int rc = d->allgatherv(mine.data(), mine.size() * 8, all, sizes);
  if (rc) return rc;
  struct Piece { uint64_t lo, hi, from, avail, n; const uint64_t* end; uint64_t nreal; };
  std::vector<Piece> pc(W);
  size_t off = 0;
  for (int r = 0; r < W; ++r) {
    if (sizes[r] < 40 || off + sizes[r] > all.size()) return d->fail(ZQ_E_ARG, "malformed piece description");
    const uint64_t* p = (const uint64_t*)(all.data() + off);   // (offsets are multiples of 8: every rank sends whole words)
    pc[r] = Piece{p[0], p[1], p[2], p[3], p[4], p + 5, p[4]};
    if (sizes[r] != (5 + p[4]) * 8) return d->fail(ZQ_E_ARG, "malformed piece description");
    // a last fragment cut short by the end of the piece is not a boundary of the stream
    if (pc[r].n && pc[r].avail < stream_total) pc[r].nreal = pc[r].n - 1;
    off += (size_t)sizes[r];
  }
  if (pc[0].lo != 0 || pc[0].from != 0 || pc[W - 1].hi != stream_total || pc[W - 1].avail != stream_total)
    return d->fail(ZQ_E_ARG, "the pieces do not cover the stream");
  for (int r = 1; r < W; ++r) if (pc[r].lo != pc[r - 1].hi) return d->fail(ZQ_E_ARG, "the pieces do not cover the stream");
  // s[r]: where rank r's chain becomes the real one
  std::vector<uint64_t> s(W + 1, 0);
  s[W] = stream_total;
  int unresolved = -1;
  uint64_t restart = 0;
  for (int r = 1; r < W && unresolved < 0; ++r) {
    const Piece& L = pc[r - 1]; const Piece& R = pc[r];
    // real boundaries of the left chain at or past lo_r: its ends >= max(s[r-1], lo_r) (a chain that starts at a real
    // boundary exactly at lo_r counts too)
    const uint64_t* lb = std::lower_bound(L.end, L.end + L.nreal, std::max(s[r - 1], R.lo));
    const uint64_t* le = L.end + L.nreal;
    bool found = false, have_first = false;
    uint64_t first_real = 0;
    if (s[r - 1] >= R.lo) { have_first = true; first_real = s[r - 1]; if (s[r - 1] == R.from || std::binary_search(R.end, R.end + R.nreal, s[r - 1])) { s[r] = s[r - 1]; found = true; } }
    for (const uint64_t* e = lb; e < le && !found; ++e) {
      if (!have_first) { have_first = true; first_real = *e; }
      if (*e == R.from || std::binary_search(R.end, R.end + R.nreal, *e)) { s[r] = *e; found = true; }
    }
    if (!found) {
      if (!have_first) return d->fail(ZQ_E_UNSUPPORTED, "overlap shorter than one fragment: the left piece holds no boundary past its end");
      unresolved = r; restart = first_real;
    }
  }
  if (unresolved >= 0) {
    out->again = 1;
    if (d->rank == unresolved) { out->restart = 1; out->restart_at = restart; }
    return ZQ_OK;
  }
  // kept fragments of every rank: those starting in [s[r], s[r+1]); fragment k starts at (k ? end[k-1] : from)
  uint64_t gtotal = 0;
  for (int r = 0; r < W; ++r) {
    const Piece& P = pc[r];
    if (s[r] < P.from) return d->fail(ZQ_E_ARG, "a piece starts past the boundary it was to join");
    auto first_from = [&](uint64_t x) -> uint64_t {      // first fragment starting at or after x (x: its `from` or one of its ends)
      if (x <= P.from) return 0;
      return std::min<uint64_t>(P.n, (uint64_t)(std::lower_bound(P.end, P.end + P.n, x) - P.end) + 1);
    };
    const uint64_t k0 = first_from(s[r]);
    const uint64_t k1 = std::max(k0, r == W - 1 ? P.n : first_from(s[r + 1]));
    if (r == d->rank) { out->first_keep = k0; out->n_keep = k1 - k0; out->global_first = gtotal; out->begin = s[r]; out->end = s[r + 1]; }
    gtotal += k1 - k0;
  }
  out->global_total = gtotal;
  return ZQ_OK;
}

}  // extern "C"
