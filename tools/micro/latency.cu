// Dependent-chain latencies of the operations the beam kernel's frame is made of, on the SM it runs on.
// One CTA of NW warps; every warp runs the chain, lane 0 of warp 0 reports (clock64 around N dependent ops).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency tools/micro/latency.cu && ./latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 256;

template <int OP>
__global__ void chain(long long *out, double seed_d, float seed_f, int seed_i) {
  __shared__ int sm[1024];
  __shared__ int hist[256];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 33 + 1) & 1023;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  double d = seed_d + threadIdx.x * 1e-9;
  float f = seed_f + threadIdx.x * 1e-6f;
  int v = seed_i + threadIdx.x;
  unsigned u = (unsigned)v;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) d = __fma_rn(d, 1.0000001, 1e-9);                       // DFMA
    if (OP == 1) d = __dadd_rn(d, 1e-9);                                  // DADD
    if (OP == 2) { f = (float)d; d = (double)f + 1e-9; }                  // F2F down + F2F up + DADD
    if (OP == 3) f = __fadd_rn(f, 1e-6f);                                 // FADD
    if (OP == 4) v = sm[v & 1023];                                        // LDS pointer chase
    if (OP == 5) v = atomicAdd(&hist[v & 255], 1) + i;                    // ATOMS with result
    if (OP == 6) v = __shfl_sync(0xffffffffu, v, (v + 1) & 31) + 1;       // SHFL
    if (OP == 7) u = __ballot_sync(0xffffffffu, (u >> (i & 7)) & 1) + u;  // VOTE + IADD
    if (OP == 8) u = __reduce_max_sync(0xffffffffu, u) + threadIdx.x;     // REDUX
    if (OP == 9) { __syncthreads(); v += 1; }                             // BAR.SYNC
    if (OP == 10) u = __fns(u | 1u, 0, 1 + (i & 3)) + u;                  // find n-th set bit
    if (OP == 11) u = __popc(u) + u;                                      // POPC
    if (OP == 12) { sm[(v & 511) + 512] = v; v = sm[(v & 511) + 512] + 1; }  // STS -> LDS same address
    if (OP == 13) { d = (double)v; v = (int)d + 1; }                      // I2F.F64 + F2I
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (d == 1.2345 || f == 1.2345f || v == -77 || u == 0xdeadbeefu) out[1] = 1;  // keep the chains alive
}

template <int OP>
static void run(const char *name, int nw, long long *d_out) {
  long long h[2];
  chain<OP><<<1, nw * 32>>>(d_out, 1.0, 1.0f, 3);
  chain<OP><<<1, nw * 32>>>(d_out, 1.0, 1.0f, 3);
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s nw=%d  %7.1f cycles/op\n", name, nw, (double)h[0] / N);
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 16);
  for (int nw : {1, 4, 8}) {
    run<0>("DFMA", nw, d_out);
    run<1>("DADD", nw, d_out);
    run<2>("F2F.F32.F64 + F2F.F64.F32 + DADD", nw, d_out);
    run<3>("FADD", nw, d_out);
    run<4>("LDS chase", nw, d_out);
    run<5>("ATOMS (returning)", nw, d_out);
    run<6>("SHFL.IDX", nw, d_out);
    run<7>("VOTE.BALLOT + IADD", nw, d_out);
    run<8>("REDUX.MAX + IADD", nw, d_out);
    run<9>("BAR.SYNC", nw, d_out);
    run<10>("__fns + IADD", nw, d_out);
    run<11>("POPC + IADD", nw, d_out);
    run<12>("STS -> LDS same address", nw, d_out);
    run<13>("I2F.F64 + F2I.F64", nw, d_out);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
