// Host-side executor of the end-to-end training pipeline (the batch feeder).
//
// Engine.train_batches issues, per step, one host->device copy of a packed batch, the structure pass
// and the step (two CUDA-graph launches), one device->host read-back and ~8 event operations.  From
// Python that is ~20 calls into torch per step, about 55 us - more than the GPU needs for the step
// (28 us) and more than the PCIe transfer of the batch (39 us).  drgnn_feed_run replays the same
// schedule from C: the per-step host cost drops to a handful of runtime calls.
//
//   copy stream  : [wait stage_free[s]]  memcpy H2D batch i -> staging slot s          record copied[s]
//   prep stream p: wait copied[s] [wait slot_free[s]]  launch structure-pass graph     record ready[s]
//   main stream  : wait ready[s]  launch step graph  D2D copy of [loss | predictions] into ring slot r
//                                                                                      record slot_free[s], stage_free[s], out[r]
//   read stream  : wait out[r]  memcpy D2H ring slot r -> pinned host row i            record read[r]
//   (the read-back is kept off the main stream: a small D2H copy costs ~8 us of latency there, a third of a step)
//
// Streams and graph-exec handles come from the caller (torch.cuda.Stream.cuda_stream,
// torch.cuda.CUDAGraph.raw_cuda_graph_exec()): runtime handles are driver handles, so they are valid
// in this library's runtime instance.  Events are owned by the library (created once per thread).
#include "common.cuh"

namespace drgnn {
static constexpr int FEED_MAX_SLOTS = 8;
struct FeedEvents {
  cudaEvent_t stage_free[FEED_MAX_SLOTS], copied[FEED_MAX_SLOTS], slot_free[FEED_MAX_SLOTS], ready[FEED_MAX_SLOTS];
  cudaEvent_t out[FEED_MAX_SLOTS], read[FEED_MAX_SLOTS];
  cudaEvent_t fork;
  bool made = false;
};
}  // namespace drgnn

using namespace drgnn;

// DRGNN_FEED_TRACE=1: time stamps (CUDA events) around the H2D copy, the structure pass and the step of the first
// 64 steps of a run, printed to stderr after a final synchronisation (diagnostic; a traced run is not a benchmark)
static constexpr int FEED_TRACE_N = 64;
struct FeedTrace {
  cudaEvent_t t0, c0[FEED_TRACE_N], c1[FEED_TRACE_N], p0[FEED_TRACE_N], p1[FEED_TRACE_N], s0[FEED_TRACE_N], s1[FEED_TRACE_N];
  bool made = false;
};

extern "C" int drgnn_feed_run(const drgnn_feed_step* steps, int32_t n, int32_t n_slots, void* main_stream, void* copy_stream,
                              void* prep_stream0, void* prep_stream1, void* read_stream, void* ring, int64_t ring_stride,
                              int32_t ring_slots) {
  DRGNN_REQUIRE(steps != nullptr || n == 0, "feed_run: steps is NULL");
  DRGNN_REQUIRE(n >= 0 && n_slots >= 1 && n_slots <= FEED_MAX_SLOTS, "feed_run: bad sizes (n %d, slots %d)", n, n_slots);
  DRGNN_REQUIRE(ring_slots >= 0 && ring_slots <= FEED_MAX_SLOTS && (ring_slots == 0 || (ring && read_stream && ring_stride > 0)),
                "feed_run: bad read-back ring");
  static thread_local FeedEvents ev;
  if (!ev.made) {
    for (int i = 0; i < FEED_MAX_SLOTS; ++i) {
      DRGNN_CHECK_CUDA(cudaEventCreateWithFlags(&ev.stage_free[i], cudaEventDisableTiming));
      DRGNN_CHECK_CUDA(cudaEventCreateWithFlags(&ev.copied[i], cudaEventDisableTiming));
      DRGNN_CHECK_CUDA(cudaEventCreateWithFlags(&ev.slot_free[i], cudaEventDisableTiming));
      DRGNN_CHECK_CUDA(cudaEventCreateWithFlags(&ev.ready[i], cudaEventDisableTiming));
      DRGNN_CHECK_CUDA(cudaEventCreateWithFlags(&ev.out[i], cudaEventDisableTiming));
      DRGNN_CHECK_CUDA(cudaEventCreateWithFlags(&ev.read[i], cudaEventDisableTiming));
    }
    DRGNN_CHECK_CUDA(cudaEventCreateWithFlags(&ev.fork, cudaEventDisableTiming));
    ev.made = true;
  }
  static thread_local FeedTrace tr;
  static const bool trace_on = [] { const char* e = getenv("DRGNN_FEED_TRACE"); return e && e[0] == '1'; }();
  const bool trace = trace_on && n >= FEED_TRACE_N;
  if (trace && !tr.made) {
    cudaEventCreate(&tr.t0);
    for (int i = 0; i < FEED_TRACE_N; ++i) {
      cudaEventCreate(&tr.c0[i]); cudaEventCreate(&tr.c1[i]); cudaEventCreate(&tr.p0[i]); cudaEventCreate(&tr.p1[i]);
      cudaEventCreate(&tr.s0[i]); cudaEventCreate(&tr.s1[i]);
    }
    tr.made = true;
  }
  cudaStream_t mainS = (cudaStream_t)main_stream, copyS = (cudaStream_t)copy_stream;
  cudaStream_t prepS[2] = {(cudaStream_t)prep_stream0, (cudaStream_t)prep_stream1};
  cudaStream_t readS = (cudaStream_t)read_stream;
  // the side streams start behind everything already queued on the main stream
  if (trace) cudaEventRecord(tr.t0, mainS);
  DRGNN_CHECK_CUDA(cudaEventRecord(ev.fork, mainS));
  DRGNN_CHECK_CUDA(cudaStreamWaitEvent(copyS, ev.fork, 0));
  DRGNN_CHECK_CUDA(cudaStreamWaitEvent(prepS[0], ev.fork, 0));
  DRGNN_CHECK_CUDA(cudaStreamWaitEvent(prepS[1], ev.fork, 0));
  if (ring_slots) DRGNN_CHECK_CUDA(cudaStreamWaitEvent(readS, ev.fork, 0));
  for (int i = 0; i < n; ++i) {
    const drgnn_feed_step& s = steps[i];
    const int slot = s.slot;
    DRGNN_REQUIRE(slot >= 0 && slot < n_slots, "feed_run: step %d uses slot %d of %d", i, slot, n_slots);
    DRGNN_REQUIRE(s.h_src && s.d_dst && s.nbytes > 0 && s.prep_graph && s.step_graph, "feed_run: step %d is incomplete", i);
    cudaStream_t ps = prepS[i & 1];
    if (i >= n_slots) DRGNN_CHECK_CUDA(cudaStreamWaitEvent(copyS, ev.stage_free[slot], 0));
    const bool tr_i = trace && i < FEED_TRACE_N;
    if (tr_i) cudaEventRecord(tr.c0[i], copyS);
    DRGNN_CHECK_CUDA(cudaMemcpyAsync(s.d_dst, s.h_src, (size_t)s.nbytes, cudaMemcpyHostToDevice, copyS));
    if (tr_i) cudaEventRecord(tr.c1[i], copyS);
    DRGNN_CHECK_CUDA(cudaEventRecord(ev.copied[slot], copyS));
    DRGNN_CHECK_CUDA(cudaStreamWaitEvent(ps, ev.copied[slot], 0));
    if (i >= n_slots) DRGNN_CHECK_CUDA(cudaStreamWaitEvent(ps, ev.slot_free[slot], 0));
    if (tr_i) cudaEventRecord(tr.p0[i], ps);
    DRGNN_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)s.prep_graph, ps));
    if (tr_i) cudaEventRecord(tr.p1[i], ps);
    DRGNN_CHECK_CUDA(cudaEventRecord(ev.ready[slot], ps));
    DRGNN_CHECK_CUDA(cudaStreamWaitEvent(mainS, ev.ready[slot], 0));
    if (tr_i) cudaEventRecord(tr.s0[i], mainS);
    DRGNN_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)s.step_graph, mainS));
    if (tr_i) cudaEventRecord(tr.s1[i], mainS);
    const bool rb = s.h_out && s.d_out && s.out_bytes > 0;
    if (rb && ring_slots) {
      DRGNN_REQUIRE(s.out_bytes <= ring_stride, "feed_run: read-back of %lld bytes exceeds the ring stride", (long long)s.out_bytes);
      const int r = i % ring_slots;
      char* rslot = reinterpret_cast<char*>(ring) + (int64_t)r * ring_stride;
      if (i >= ring_slots) DRGNN_CHECK_CUDA(cudaStreamWaitEvent(mainS, ev.read[r], 0));
      DRGNN_CHECK_CUDA(cudaMemcpyAsync(rslot, s.d_out, (size_t)s.out_bytes, cudaMemcpyDeviceToDevice, mainS));
      DRGNN_CHECK_CUDA(cudaEventRecord(ev.out[r], mainS));
      DRGNN_CHECK_CUDA(cudaStreamWaitEvent(readS, ev.out[r], 0));
      DRGNN_CHECK_CUDA(cudaMemcpyAsync(s.h_out, rslot, (size_t)s.out_bytes, cudaMemcpyDeviceToHost, readS));
      DRGNN_CHECK_CUDA(cudaEventRecord(ev.read[r], readS));
    } else if (rb) {
      DRGNN_CHECK_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)s.out_bytes, cudaMemcpyDeviceToHost, mainS));
    }
    DRGNN_CHECK_CUDA(cudaEventRecord(ev.slot_free[slot], mainS));
    DRGNN_CHECK_CUDA(cudaEventRecord(ev.stage_free[slot], mainS));
  }
  // join: later work on the main stream also follows the side streams
  if (n > 0) {
    DRGNN_CHECK_CUDA(cudaEventRecord(ev.fork, copyS));
    DRGNN_CHECK_CUDA(cudaStreamWaitEvent(mainS, ev.fork, 0));
    for (int k = 0; k < 2; ++k) {
      DRGNN_CHECK_CUDA(cudaEventRecord(ev.fork, prepS[k]));
      DRGNN_CHECK_CUDA(cudaStreamWaitEvent(mainS, ev.fork, 0));
    }
    if (ring_slots) {
      DRGNN_CHECK_CUDA(cudaEventRecord(ev.fork, readS));
      DRGNN_CHECK_CUDA(cudaStreamWaitEvent(mainS, ev.fork, 0));
    }
  }
  if (trace) {
    cudaDeviceSynchronize();
    fprintf(stderr, "feed trace (us since the start of the run): step | copy start end | structure start end | step start end\n");
    for (int i = 0; i < FEED_TRACE_N; ++i) {
      float a, b, c, d, e, f;
      cudaEventElapsedTime(&a, tr.t0, tr.c0[i]); cudaEventElapsedTime(&b, tr.t0, tr.c1[i]);
      cudaEventElapsedTime(&c, tr.t0, tr.p0[i]); cudaEventElapsedTime(&d, tr.t0, tr.p1[i]);
      cudaEventElapsedTime(&e, tr.t0, tr.s0[i]); cudaEventElapsedTime(&f, tr.t0, tr.s1[i]);
      fprintf(stderr, "  %2d | %8.1f %8.1f | %8.1f %8.1f | %8.1f %8.1f\n", i, 1e3 * a, 1e3 * b, 1e3 * c, 1e3 * d, 1e3 * e, 1e3 * f);
    }
  }
  return DRGNN_OK;
}
