// model.fit's inner loop for NeuMF (src/models/RModel.py:130-137 over the batches of bootstrapDataset,
// src/models/NeuMFModel.py:102-123) as ONE C call: for every listed batch the fused forward/backward and the
// optimizer step are enqueued back to back by this thread; nothing synchronises, losses land in device memory.
// Measured (profiles/neumf_loop_case.py): this does NOT make small batches faster -- 76 us/step at the reference's
// batch of 128 either way: the step is bound by the latency of its six dependent kernels (one or two CTAs each at
// that size), not by the Python that used to sit between them.  It is the epoch-level entry point all the same:
// one call per epoch, no per-step host work to overlap with.
#include "common.cuh"

int brk_neumf_step_fused(brk_ctx* ctx, const brk_neumf_model* m, const brk_neumf_shards* sh, const int32_t* u,
                         const int32_t* i, const float* y, int64_t batch, int64_t global_batch, int64_t first_index,
                         int32_t training, uint32_t dropout_seed, uint32_t dropout_epoch, const brk_neumf_workspace* ws,
                         float* out, float* loss_out, cudaStream_t st, int* rc_out, int* handled,
                         const brk_adam_hyper* adam_h, int64_t* adam_state);

// One training step = fused forward/backward + optimizer.  With the stock Keras Adam and a model the one-launch
// tensor-core kernel covers (csrc/neumf_fused.cu: tensor_cores = 1 class spec numFactor 32 / 64, or the He et al.
// variant) and a batch that fits on chip, the Adam pass over the four tables and the dense block runs INSIDE that
// launch, after its last grid barrier: one kernel per step.  Otherwise brk_neumf_step + the optimizer kernels.
extern "C" int brk_neumf_train_step(brk_ctx* ctx, const brk_neumf_model* m, const int32_t* u, const int32_t* i,
                                    const float* y, int64_t batch, int64_t first_index, uint32_t dropout_seed,
                                    uint32_t dropout_epoch, brk_adam_hyper h, int64_t* adam_state, int32_t lazy_adam,
                                    const brk_neumf_workspace* ws, float* out, float* loss_out, void* stream) {
  BRK_REQUIRE(ctx && m && u && i && y && ws && out && adam_state, BRK_E_ARG, "brk_neumf_train_step: null argument");
  BRK_REQUIRE(batch > 0, BRK_E_ARG, "brk_neumf_train_step: batch=%lld", (long long)batch);
  const bool variant = (m->EMF > 0 && m->EMF != m->E) || m->mf_mode != 0 || m->no_batch_norm != 0;
  if (!lazy_adam && (m->tensor_cores || variant) && m->uMLP.w && m->dense.w && m->dense.g && m->bn_moving && ws->acc) {
    int rc2 = 0, handled = 0;
    brk_neumf_step_fused(ctx, m, nullptr, u, i, y, batch, 0, first_index, 1, dropout_seed, dropout_epoch, ws, out, loss_out,
                         (cudaStream_t)stream, &rc2, &handled, &h, adam_state);
    if (handled) return rc2;
  }
  int rc = brk_neumf_step(ctx, m, u, i, y, batch, 0, first_index, 1, dropout_seed, dropout_epoch, ws, out, loss_out, stream);
  if (rc) return rc;
  const brk_table all[5] = {m->uMLP, m->iMLP, m->uMF, m->iMF, m->dense};
  if (!lazy_adam) return brk_adam_dense_keras(ctx, all, 5, h, adam_state, 1, stream);   // exact Keras: every element, tables included
  rc = brk_adam_dense_keras(ctx, all + 4, 1, h, adam_state, 0, stream);
  if (!rc) rc = brk_adam_rows(ctx, all, 4, h, adam_state, 1, stream);
  return rc;
}

extern "C" int brk_neumf_train_steps(brk_ctx* ctx, const brk_neumf_model* m, const int32_t* u, const int32_t* i,
                                     const float* y, int64_t n_rows, int64_t batch, const int64_t* batch_index_host,
                                     int32_t n_steps, uint32_t dropout_seed, uint32_t dropout_epoch,
                                     brk_adam_hyper h, int64_t* adam_state, int32_t lazy_adam,
                                     const brk_neumf_workspace* ws, float* out, float* losses, void* stream) {
  BRK_REQUIRE(ctx && m && u && i && y && ws && out && adam_state, BRK_E_ARG, "brk_neumf_train_steps: null argument");
  BRK_REQUIRE(n_rows > 0 && batch > 0 && n_steps >= 0, BRK_E_ARG, "brk_neumf_train_steps: n_rows=%lld batch=%lld n_steps=%d",
              (long long)n_rows, (long long)batch, n_steps);
  BRK_REQUIRE(n_steps == 0 || batch_index_host, BRK_E_ARG, "brk_neumf_train_steps: batch_index_host is null");
  const int64_t n_batches = (n_rows + batch - 1) / batch;
  for (int32_t s = 0; s < n_steps; ++s) {
    const int64_t b = batch_index_host[s];
    BRK_REQUIRE(b >= 0 && b < n_batches, BRK_E_ARG, "brk_neumf_train_steps: batch index %lld of %lld", (long long)b,
                (long long)n_batches);
    const int64_t off = b * batch;
    const int64_t count = (n_rows - off < batch) ? n_rows - off : batch;          // ragged last batch
    const int rc = brk_neumf_train_step(ctx, m, u + off, i + off, y + off, count, off, dropout_seed, dropout_epoch, h, adam_state,
                                        lazy_adam, ws, out, losses ? losses + s : nullptr, stream);
    if (rc) return rc;
  }
  return 0;
}

// Host-fed training steps: the frame lives in page-locked HOST memory in the loader's batch-major layout
// packed_host [n_batches][3][batch] (user ids, item ids, labels as float bits -- what a loader producing
// ({"user": ids, "item": ids}, label) batches writes, NeuMFModel.py:111-117).  Per step ONE cudaMemcpyAsync of that
// batch's 12 * batch bytes on the context's copy stream into one of four staging slots, then brk_neumf_train_step on
// the slot; the copy of step s + 1..s + 3 overlaps the compute of step s.  The step losses return in ONE D2H copy into
// losses_host (valid after the stream is synchronised).  Full batches only.
extern "C" int64_t brk_neumf_host_stage_ints(int64_t batch) { return 4 * 3 * batch; }

extern "C" int brk_neumf_train_steps_host(brk_ctx* ctx, const brk_neumf_model* m, const int32_t* packed_host, int64_t n_batches,
                                          int64_t batch, const int64_t* batch_index_host, int32_t n_steps, uint32_t dropout_seed,
                                          uint32_t dropout_epoch, brk_adam_hyper h, int64_t* adam_state, int32_t lazy_adam,
                                          const brk_neumf_workspace* ws, int32_t* d_stage, float* out, float* d_losses,
                                          float* losses_host, void* stream) {
  BRK_REQUIRE(ctx && m && packed_host && ws && d_stage && out && d_losses && adam_state, BRK_E_ARG,
              "brk_neumf_train_steps_host: null argument");
  BRK_REQUIRE(n_batches > 0 && batch > 0 && n_steps >= 0 && (n_steps == 0 || batch_index_host), BRK_E_ARG,
              "brk_neumf_train_steps_host: n_batches=%lld batch=%lld n_steps=%d", (long long)n_batches, (long long)batch, n_steps);
  if (n_steps == 0) return 0;
  if (int rc = brk_ctx_ensure_copy(ctx)) return rc;
  cudaStream_t st = (cudaStream_t)stream, cs = ctx->copy_stream;
  constexpr int kSlots = BRK_STAGE_EVENTS;
  const size_t slot_ints = size_t(3) * size_t(batch);
  // the copy stream starts after whatever the caller has queued on `st` (earlier steps may still read the slots)
  BRK_CUDA(cudaEventRecord(ctx->ev_go, st));
  BRK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_go, 0));
  for (int32_t s = 0; s < n_steps; ++s) {
    const int64_t b = batch_index_host[s];
    BRK_REQUIRE(b >= 0 && b < n_batches, BRK_E_ARG, "brk_neumf_train_steps_host: batch index %lld of %lld", (long long)b,
                (long long)n_batches);
    const int q = s % kSlots;
    int32_t* slot = d_stage + size_t(q) * slot_ints;
    if (s >= kSlots) BRK_CUDA(cudaStreamWaitEvent(cs, ctx->ev_done[q], 0));      // the step that last used this slot is done
    BRK_CUDA(cudaMemcpyAsync(slot, packed_host + size_t(b) * slot_ints, slot_ints * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
    BRK_CUDA(cudaEventRecord(ctx->ev_ready[q], cs));
    BRK_CUDA(cudaStreamWaitEvent(st, ctx->ev_ready[q], 0));
    const int rc = brk_neumf_train_step(ctx, m, slot, slot + batch, reinterpret_cast<const float*>(slot + 2 * batch), batch,
                                        b * batch, dropout_seed, dropout_epoch, h, adam_state, lazy_adam, ws, out, d_losses + s, stream);
    if (rc) return rc;
    BRK_CUDA(cudaEventRecord(ctx->ev_done[q], st));
  }
  if (losses_host) BRK_CUDA(cudaMemcpyAsync(losses_host, d_losses, size_t(n_steps) * sizeof(float), cudaMemcpyDeviceToHost, st));
  return 0;
}
