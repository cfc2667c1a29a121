// "The step after" the path: the reference hands back dense [B, beam, T] int32 tensors of
// which only [b, p, :out_len[b, p]] is meaningful (reference binding.cpp:79-99, README.md:99-109) -- at config 2
// about a fifth of the columns.  These two kernels compact the decoded rows on the device into a ragged (CSR)
// layout, so that a caller that wants the results on the host moves sum(out_len) instead of B * beam * T elements:
//   offsets[r]            int64, r = b * beam + p, exclusive prefix sum of the row lengths (rows p >= n_results[b]
//                         have length 0); offsets[B * beam] = total
//   packed_tokens / packed_timesteps [total]
#pragma once
#include <cstdint>

namespace ctc {

// single CTA: exclusive scan of the row lengths (B * beam is at most a few hundred thousand)
__global__ void __launch_bounds__(1024) pack_offsets_kernel(const int *lens, const int *n_results, int B, int K,
                                                            long long *offsets) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long rows = (long long)B * K;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < rows; base += 1024) {
    const long long r = base + tid;
    long long v = 0;
    if (r < rows) {
      const int b = (int)(r / K), p = (int)(r - (long long)b * K);
      v = p < n_results[b] ? lens[r] : 0;
    }
    long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long o = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += o;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += o;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const long long before = s_carry + (warp > 0 ? s_warp[warp - 1] : 0) + (x - v);
    if (r < rows) offsets[r] = before;
    __syncthreads();
    if (tid == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (tid == 0) offsets[rows] = s_carry;
}

// one warp per row: copy the meaningful prefix of the dense row to its place in the packed arrays
__global__ void __launch_bounds__(256) pack_rows_kernel(const int *tokens, const int *timesteps, const int *lens,
                                                        const int *n_results, int B, int K, int T,
                                                        const long long *offsets, int *packed_tokens,
                                                        int *packed_timesteps, long long capacity) {
  const long long rows = (long long)B * K;
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(r / K), p = (int)(r - (long long)b * K);
    if (p >= n_results[b]) continue;
    const int len = lens[r];
    const long long off = offsets[r];
    if (off + len > capacity) continue;  // caller's buffers too small: offsets[rows] tells
    for (int i = lane; i < len; i += 32) {
      packed_tokens[off + i] = tokens[r * T + i];
      packed_timesteps[off + i] = timesteps[r * T + i];
    }
  }
}

}  // namespace ctc
