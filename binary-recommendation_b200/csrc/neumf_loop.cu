// model.fit's inner loop for NeuMF (src/models/RModel.py:130-137 over the batches of bootstrapDataset,
// src/models/NeuMFModel.py:102-123) as ONE C call: for every listed batch the fused forward/backward and the
// optimizer step are enqueued back to back by this thread; nothing synchronises, losses land in device memory.
// Measured (profiles/neumf_loop_case.py): this does NOT make small batches faster -- 76 us/step at the reference's
// batch of 128 either way: the step is bound by the latency of its six dependent kernels (one or two CTAs each at
// that size), not by the Python that used to sit between them.  It is the epoch-level entry point all the same:
// one call per epoch, no per-step host work to overlap with.
#include "common.cuh"

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
  const brk_table all[5] = {m->uMLP, m->iMLP, m->uMF, m->iMF, m->dense};
  for (int32_t s = 0; s < n_steps; ++s) {
    const int64_t b = batch_index_host[s];
    BRK_REQUIRE(b >= 0 && b < n_batches, BRK_E_ARG, "brk_neumf_train_steps: batch index %lld of %lld", (long long)b,
                (long long)n_batches);
    const int64_t off = b * batch;
    const int64_t count = (n_rows - off < batch) ? n_rows - off : batch;          // ragged last batch
    int rc = brk_neumf_step(ctx, m, u + off, i + off, y + off, count, 0, off, 1, dropout_seed, dropout_epoch, ws, out,
                            losses ? losses + s : nullptr, stream);
    if (rc) return rc;
    if (!lazy_adam) {
      rc = brk_adam_dense_keras(ctx, all, 5, h, adam_state, 1, stream);           // exact Keras: every element, tables included
    } else {
      rc = brk_adam_dense_keras(ctx, all + 4, 1, h, adam_state, 0, stream);
      if (!rc) rc = brk_adam_rows(ctx, all, 4, h, adam_state, 1, stream);
    }
    if (rc) return rc;
  }
  return 0;
}
