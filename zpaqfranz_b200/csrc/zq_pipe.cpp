// zq_pipe.cpp -- batches in flight: the GPU-side counterpart of CompressJob's queue of blocks between the reader
// and the compressor threads (CompressJob::appendz / compressThread / writeThread, Z:71364-71520).
//
// One batch (zq_compress_blocks*) ends with a tail: the last units of the LZ77 parse run on a few warps while the
// rest of the machine idles, and the host-pointer form also has its H2D copy in front and its D2H copy behind.
// A front end that keeps reading while blocks compress has more than one batch outstanding; this file gives it
// `depth` lanes, each a worker thread with its own context and stream, so batch k+1's copies and kernels fill
// the gaps of batch k.  Results are identical to the synchronous calls (same code underneath).
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <map>
#include <set>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zq_b200.h"

namespace {
struct Job {
  int ticket = 0, n = 0;
  const uint8_t* in = nullptr; const uint64_t* in_off = nullptr; const uint32_t* in_len = nullptr;
  const char* const* method = nullptr; const char* const* filename = nullptr; const char* const* comment = nullptr;
  int uniform = 0, dosha1 = 1, device_pointers = 0;
  uint8_t* out = nullptr; uint64_t out_cap = 0; uint64_t* out_off = nullptr; uint32_t* out_len = nullptr;
};
struct Done { int rc; std::string err; };
}  // namespace

struct zq_pipe {
  std::vector<zq_ctx*> ctx;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::mutex compute;        // host-pointer batches: one batch's kernels at a time, the others copy meanwhile
  std::mutex copy_in;        // and one input copy at a time
  bool gated = true;
  std::condition_variable cv_job, cv_done;
  std::deque<Job> queue;
  std::map<int, Done> done;
  std::set<int> collected;   // tickets already waited for
  int next_ticket = 0;
  bool stop = false;
  std::string last_error;

  static void gate(void* self, int acquire) {
    zq_pipe* p = static_cast<zq_pipe*>(self);
    switch (acquire) {
      case 1: p->compute.lock(); break;
      case 0: p->compute.unlock(); break;
      case 2: p->copy_in.lock(); break;
      default: p->copy_in.unlock(); break;
    }
  }

  void run(int lane) {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return stop || !queue.empty(); });
        if (queue.empty()) return;
        j = queue.front(); queue.pop_front();
      }
      zq_ctx* c = ctx[lane];
      zq_set_compute_gate(c, (gated && !j.device_pointers) ? &zq_pipe::gate : nullptr, this);
      const int rc = j.device_pointers
          ? zq_compress_blocks_device(c, j.n, j.in, j.in_off, j.in_len, j.method, j.filename, j.comment, j.uniform, j.dosha1, j.out, j.out_cap, j.out_off, j.out_len)
          : zq_compress_blocks(c, j.n, j.in, j.in_off, j.in_len, j.method, j.filename, j.comment, j.uniform, j.dosha1, j.out, j.out_cap, j.out_off, j.out_len);
      {
        std::lock_guard<std::mutex> lk(mu);
        done[j.ticket] = Done{rc, rc ? zq_last_error(c) : ""};
      }
      cv_done.notify_all();
    }
  }
};

extern "C" {

zq_pipe* zq_pipe_create(int device, int depth) {
  if (depth < 1) depth = 1;
  if (depth > 8) depth = 8;
  zq_pipe* p = new zq_pipe();
  if (const char* s = getenv("ZQ_PIPE_GATE")) p->gated = atoi(s) != 0;
  for (int i = 0; i < depth; ++i) {
    zq_ctx* c = zq_create(device);
    if (!c) { for (zq_ctx* x : p->ctx) zq_destroy(x); delete p; return nullptr; }
    p->ctx.push_back(c);
  }
  for (int i = 0; i < depth; ++i) p->workers.emplace_back([p, i] { p->run(i); });
  return p;
}

void zq_pipe_destroy(zq_pipe* p) {
  if (!p) return;
  { std::lock_guard<std::mutex> lk(p->mu); p->stop = true; }
  p->cv_job.notify_all();
  for (auto& t : p->workers) t.join();
  for (zq_ctx* c : p->ctx) zq_destroy(c);
  delete p;
}

int zq_pipe_submit(zq_pipe* p, int n, const uint8_t* in_base, const uint64_t* in_off, const uint32_t* in_len,
                   const char* const* method, const char* const* filename, const char* const* comment, int uniform, int dosha1,
                   int device_pointers, uint8_t* out_base, uint64_t out_cap, uint64_t* out_off, uint32_t* out_len) {
  if (!p) return ZQ_E_NODEVICE;
  Job j;
  j.n = n; j.in = in_base; j.in_off = in_off; j.in_len = in_len; j.method = method; j.filename = filename; j.comment = comment;
  j.uniform = uniform; j.dosha1 = dosha1; j.device_pointers = device_pointers; j.out = out_base; j.out_cap = out_cap;
  j.out_off = out_off; j.out_len = out_len;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    j.ticket = p->next_ticket++;
    p->queue.push_back(j);
  }
  p->cv_job.notify_one();
  return j.ticket;
}

int zq_pipe_wait(zq_pipe* p, int ticket) {
  if (!p) return ZQ_E_NODEVICE;
  std::unique_lock<std::mutex> lk(p->mu);
  if (ticket < 0 || ticket >= p->next_ticket || p->collected.count(ticket)) return ZQ_E_ARG;   // unknown, or waited for already
  p->cv_done.wait(lk, [&] { return p->done.count(ticket) != 0; });
  Done d = p->done[ticket];
  p->done.erase(ticket);
  p->collected.insert(ticket);
  if (d.rc) p->last_error = d.err;
  return d.rc;
}

const char* zq_pipe_last_error(zq_pipe* p) { return p ? p->last_error.c_str() : "no pipe"; }

uint64_t zq_pipe_launch_count(zq_pipe* p) {
  uint64_t s = 0;
  if (p) for (zq_ctx* c : p->ctx) s += zq_launch_count(c);
  return s;
}

}  // extern "C"
