// bk_loop.cuh — the device-resident iteration loop driver shared by CG / BiCGStab / GMRES.
//
// The reference decides `if k >= maxiter or rs <= atol2: break` with a Python bool() on a device
// tensor — one host<->device sync per iteration (torch_sparse_linalg.py:841, :895-936, :798).
// Here the stop test is evaluated by the epilogue of the kernel that finishes the residual
// reduction; it sets bk_dev_state.done, and every later kernel of the sequence begins with
// `if (st->done) return;`, i.e. is an exact no-op.  The host therefore only has to keep the
// stream fed: it enqueues `chunk` iterations at a time — as one cached CUDA graph launch
// (BK_LOOP_GRAPH) or as plain launches (BK_LOOP_STREAM) — followed by an async copy of the
// state word, and looks at the copy of chunk i only after chunk i+1 has been enqueued, so the
// GPU never waits for the host.
#pragma once

#include "bk_internal.cuh"

static __global__ void bk_state_set_kernel(bk_dev_state* st, const bk_dev_state v) { *st = v; }

// Look up / build the graph of one chunk.  `enqueue` must enqueue the chunk on the stream it is given.
template <typename F>
static int bk_chunk_graph(bk_handle* h, const uint64_t key[6], F enqueue, cudaGraphExec_t* out) {
  *out = nullptr;
  for (int i = 0; i < 8; ++i) {
    if (h->graphs[i].valid && !memcmp(h->graphs[i].key, key, sizeof(uint64_t) * 6)) {
      *out = h->graphs[i].exec;
      return BK_OK;
    }
  }
  cudaGraph_t graph = nullptr;
  BK_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue(h->cap_stream);
  cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
  if (rc != BK_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return bk_fail(BK_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return bk_fail(BK_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  }
  // insert (round-robin eviction)
  static int victim = 0;
  int slot = -1;
  for (int i = 0; i < 8; ++i)
    if (!h->graphs[i].valid) {
      slot = i;
      break;
    }
  if (slot < 0) {
    slot = (victim++) & 7;
    cudaGraphExecDestroy(h->graphs[slot].exec);
  }
  h->graphs[slot].exec = exec;
  memcpy(h->graphs[slot].key, key, sizeof(uint64_t) * 6);
  h->graphs[slot].valid = 1;
  *out = exec;
  return BK_OK;
}

// Run chunks until the device reports done.  The poll of chunk i happens after chunk i+1 is enqueued.
template <typename F>
static int bk_run_loop(bk_handle* h, cudaStream_t s, bool use_graph, const uint64_t key[6], F enqueue_chunk,
                       int64_t* chunks_out = nullptr) {
  cudaGraphExec_t exec = nullptr;
  if (use_graph) {
    int rc = bk_chunk_graph(h, key, enqueue_chunk, &exec);
    if (rc != BK_OK) {
      // graph path unavailable: say so once, keep going with plain launches (same kernels)
      static int warned = 0;
      if (!warned) {
        fprintf(stderr, "[bk_krylov] CUDA graph path disabled: %s\n", bk_last_error());
        warned = 1;
      }
      exec = nullptr;
    }
  }
  h->last_loop_mode = exec ? BK_LOOP_GRAPH : BK_LOOP_STREAM;
  int slot = 0, prev = 0;
  bool pending = false;
  int64_t chunks = 0;
  for (;;) {
    ++chunks;
    if (exec) {
      BK_CUDA(cudaGraphLaunch(exec, s));
    } else {
      BK_TRY(enqueue_chunk(s));
    }
    BK_CUDA(cudaMemcpyAsync(&h->st_host[slot], h->st, sizeof(bk_dev_state), cudaMemcpyDeviceToHost, s));
    BK_CUDA(cudaEventRecord(h->ev[slot], s));
    if (pending) {
      BK_CUDA(cudaEventSynchronize(h->ev[prev]));
      if (h->st_host[prev].done) break;
    }
    pending = true;
    prev = slot;
    slot ^= 1;
  }
  if (chunks_out) *chunks_out = chunks;
  return BK_OK;
}

static inline int bk_pick_chunk(const bk_handle* h, double bytes_per_iter, int kernels_per_iter) {
  if (h->chunk > 0) return (h->chunk + 1) & ~1;
  // aim at ~2 ms of GPU work per chunk (>> the ~20 us poll latency), assuming ~5 TB/s and ~3 us per launch; launch-bound
  // (small) systems get ~0.3 ms chunks: the loop overshoots by up to two chunks of guarded no-op launches, which cost
  // nothing next to a 300 us iteration but as much as a real iteration when that takes 10 us
  const double t_iter = bytes_per_iter / 5.0e12 + kernels_per_iter * 3.0e-6;
  const double target = (t_iter < 40.0e-6) ? 0.3e-3 : 2.0e-3;
  int c = (int)(target / t_iter);
  if (c < 2) c = 2;
  if (c > 64) c = 64;
  return (c + 1) & ~1;
}
