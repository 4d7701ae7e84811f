// Host side of the input feed (no device code): gather ragged utterances from the loader's packed frame store into one
// padded batch buffer -- normally the pinned staging memory the H2D copy reads from -- with a few threads.
// replaces: the per-utterance np.pad + np.array of pad_to_longest (U/instances_handler.py:118-139) for the features.
#include "common.cuh"
#include <string.h>
#include <thread>
#include <vector>

namespace pka {

static void pack_rows(const float* frames, const int64_t* offsets, const int64_t* idx, int row_lo, int row_hi, int t_pad,
                      int dim, float* src, uint8_t* mask) {
  for (int b = row_lo; b < row_hi; ++b) {
    const int64_t u = idx[b];
    const int64_t t = offsets[u + 1] - offsets[u];
    float* dst = src + (size_t)b * t_pad * dim;
    memcpy(dst, frames + (size_t)offsets[u] * dim, (size_t)t * dim * sizeof(float));
    memset(dst + (size_t)t * dim, 0, (size_t)(t_pad - t) * dim * sizeof(float));
    memset(mask + (size_t)b * t_pad, 1, (size_t)t);
    memset(mask + (size_t)b * t_pad + t, 0, (size_t)(t_pad - t));
  }
}

}  // namespace pka

// frames: float[total_frames, dim] (all utterances back to back); offsets: int64[n_total + 1]; idx: int64[n_utt] utterance
// numbers of this batch; src: float[n_utt, t_pad, dim], mask: uint8[n_utt, t_pad] (1 = real frame) -- both fully written.
extern "C" int pka_host_pack_batch(const float* frames, const int64_t* offsets, int64_t n_total, const int64_t* idx, int n_utt,
                                   int t_pad, int dim, float* src, uint8_t* mask, int n_threads) {
  PKA_REQUIRE(frames && offsets && idx && src && mask && n_utt > 0 && t_pad >= 0 && dim > 0, PKA_EINVAL,
              "host_pack_batch: bad arguments");
  for (int b = 0; b < n_utt; ++b) {
    PKA_REQUIRE(idx[b] >= 0 && idx[b] < n_total, PKA_EINVAL, "host_pack_batch: utterance index %lld out of range", (long long)idx[b]);
    const int64_t t = offsets[idx[b] + 1] - offsets[idx[b]];
    PKA_REQUIRE(t >= 0 && t <= t_pad, PKA_EINVAL, "host_pack_batch: utterance %lld has %lld frames, padded length is %d",
                (long long)idx[b], (long long)t, t_pad);
  }
  if (n_threads > n_utt) n_threads = n_utt;
  if (n_threads <= 1) {
    pka::pack_rows(frames, offsets, idx, 0, n_utt, t_pad, dim, src, mask);
    return 0;
  }
  std::vector<std::thread> workers;
  workers.reserve(n_threads - 1);
  const int per = (n_utt + n_threads - 1) / n_threads;
  for (int w = 1; w < n_threads; ++w) {
    const int lo = w * per, hi = lo + per < n_utt ? lo + per : n_utt;
    if (lo < hi) workers.emplace_back(pka::pack_rows, frames, offsets, idx, lo, hi, t_pad, dim, src, mask);
  }
  pka::pack_rows(frames, offsets, idx, 0, per < n_utt ? per : n_utt, t_pad, dim, src, mask);
  for (auto& th : workers) th.join();
  return 0;
}
