// Microbenchmark behind the scorer path's host<->kernel handshake: how fast can a running
// kernel exchange a flag + a small payload with host threads through device-mapped pinned memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o sysmem_pingpong sysmem_pingpong.cu && ./sysmem_pingpong
// Prints, for several CTA counts / poll styles, the mean round trip (GPU raises done -> host answers go + payload ->
// GPU has the payload in registers) in microseconds.
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct alignas(128) Line { int go; int count; int payload[30]; };   // host -> GPU, one 128-byte line per CTA
struct alignas(128) Done { int done; int pad[31]; };                // GPU -> host

// mode 0: thread 0 spins on line.go (volatile 4-byte loads), then 32 lanes read the line
// mode 1: warp 0 polls the whole line with one coalesced 128-byte load per poll
// mode 2: like 1 with __nanosleep(sleep_ns) between polls
__global__ void pingpong(Done *done, const Line *line, int rounds, int mode, int sleep_ns, long long *cycles, int *sink) {
  const int b = blockIdx.x;
  __shared__ int s_line[32];
  long long t_wait = 0;
  int acc = 0;
  for (int r = 1; r <= rounds; ++r) {
    __syncthreads();
    const long long c0 = clock64();
    if (threadIdx.x == 0) {
      __threadfence_system();
      *(volatile int *)&done[b].done = r;
    }
    if (mode == 0) {
      if (threadIdx.x == 0) {
        while (*(volatile const int *)&line[b].go < r && clock64() - c0 < 2000000000ll) { }
        __threadfence_system();
      }
      __syncthreads();
      if (threadIdx.x < 32) s_line[threadIdx.x] = ((volatile const int *)&line[b])[threadIdx.x];
    } else if (threadIdx.x < 32) {
      int v;
      for (;;) {
        v = ((volatile const int *)&line[b])[threadIdx.x];
        const int go = __shfl_sync(0xffffffffu, v, 0);
        if (go >= r || clock64() - c0 > 2000000000ll) break;
        if (mode == 2) __nanosleep(sleep_ns);
      }
      // the line is written payload first, go last (host release store): re-read once so that the payload is
      // at least as new as the flag
      v = ((volatile const int *)&line[b])[threadIdx.x];
      s_line[threadIdx.x] = v;
    }
    __syncthreads();
    acc += s_line[(threadIdx.x & 31)];
    if (threadIdx.x == 0) t_wait += clock64() - c0;
    // some "work" between handshakes so that CTAs are not in perfect lock step
    for (int k = 0; k < 200 + 37 * (b & 7); ++k) acc = acc * 1664525 + 1013904223;
  }
  if (threadIdx.x == 0) cycles[b] = t_wait;
  if (acc == 12345) sink[0] = acc;
}

// mode 3: the host answers into DEVICE memory with cudaMemcpyAsync; the kernel polls device memory
__global__ void pingpong_dev(Done *done, const Line *dline, int rounds, long long *cycles, int *sink) {
  const int b = blockIdx.x;
  __shared__ int s_line[32];
  long long t_wait = 0;
  int acc = 0;
  for (int r = 1; r <= rounds; ++r) {
    __syncthreads();
    const long long c0 = clock64();
    if (threadIdx.x == 0) {
      __threadfence_system();
      *(volatile int *)&done[b].done = r;
    }
    if (threadIdx.x < 32) {
      int v;
      for (;;) {
        v = ((volatile const int *)&dline[b])[threadIdx.x];
        if (__shfl_sync(0xffffffffu, v, 0) >= r || clock64() - c0 > 2000000000ll) break;
      }
      v = ((volatile const int *)&dline[b])[threadIdx.x];
      s_line[threadIdx.x] = v;
    }
    __syncthreads();
    acc += s_line[(threadIdx.x & 31)];
    if (threadIdx.x == 0) t_wait += clock64() - c0;
    for (int k = 0; k < 200 + 37 * (b & 7); ++k) acc = acc * 1664525 + 1013904223;
  }
  if (threadIdx.x == 0) cycles[b] = t_wait;
  if (acc == 12345) sink[0] = acc;
}

int main() {
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const int rounds = 2000;
  Done *done; Line *line, *dline;
  long long *cycles; int *sink;
  CK(cudaMallocHost(&done, sizeof(Done) * 256));
  CK(cudaMallocHost(&line, sizeof(Line) * 256));
  CK(cudaMalloc(&dline, sizeof(Line) * 256));
  CK(cudaMallocManaged(&cycles, 8 * 256));
  CK(cudaMalloc(&sink, 4));
  cudaStream_t ks;
  CK(cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking));
  struct Cfg { int ctas, threads, mode, sleep_ns; };
  std::vector<Cfg> cfgs;
  for (int ctas : {1, 8, 64, 148})
    for (int threads : {1, 4, 8})
      for (int mode : {0, 1, 2, 3}) {
        if (threads > ctas) continue;
        if (mode == 2) { cfgs.push_back({ctas, threads, 2, 500}); cfgs.push_back({ctas, threads, 2, 2000}); }
        else cfgs.push_back({ctas, threads, mode, 0});
      }
  for (const Cfg &c : cfgs) {
    memset(done, 0, sizeof(Done) * 256);
    memset(line, 0, sizeof(Line) * 256);
    CK(cudaMemset(dline, 0, sizeof(Line) * 256));
    CK(cudaDeviceSynchronize());
    std::atomic<int> stop{0};
    std::vector<std::thread> pool;
    for (int w = 0; w < c.threads; ++w)
      pool.emplace_back([&, w]() {
        cudaStream_t cs = nullptr;
        if (c.mode == 3) CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        std::vector<int> served(c.ctas, 0);
        int remaining = 0;
        for (int b = w; b < c.ctas; b += c.threads) ++remaining;
        while (remaining > 0 && !stop.load(std::memory_order_relaxed)) {
          for (int b = w; b < c.ctas; b += c.threads) {
            if (served[b] >= rounds) continue;
            const int d = reinterpret_cast<std::atomic<int> *>(&done[b].done)->load(std::memory_order_acquire);
            if (d <= served[b]) continue;
            line[b].count = 3;
            for (int k = 0; k < 6; ++k) line[b].payload[k] = d + k;
            if (c.mode == 3) {
              line[b].go = d;
              CK(cudaMemcpyAsync(&dline[b], &line[b], sizeof(Line), cudaMemcpyHostToDevice, cs));
            } else {
              reinterpret_cast<std::atomic<int> *>(&line[b].go)->store(d, std::memory_order_release);
            }
            served[b] = d;
            if (d >= rounds) --remaining;
          }
        }
        if (cs) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
      });
    const auto t0 = std::chrono::steady_clock::now();
    if (c.mode == 3) pingpong_dev<<<c.ctas, 256, 0, ks>>>(done, dline, rounds, cycles, sink);
    else pingpong<<<c.ctas, 256, 0, ks>>>(done, line, rounds, c.mode, c.sleep_ns, cycles, sink);
    CK(cudaGetLastError());
    // watchdog: never hang the box
    std::thread dog([&]() {
      for (int i = 0; i < 200 && !stop.load(); ++i) std::this_thread::sleep_for(std::chrono::milliseconds(100));
    });
    CK(cudaStreamSynchronize(ks));
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    stop.store(1);
    for (auto &t : pool) t.join();
    dog.join();
    double mean = 0;
    for (int b = 0; b < c.ctas; ++b) mean += (double)cycles[b];
    mean /= c.ctas;
    const char *names[] = {"spin4B", "warp128B", "warp128B+sleep", "memcpy->dev"};
    printf("ctas %3d host threads %d mode %-15s sleep %4d ns: round trip %.2f us (wall %.1f ms for %d rounds)\n", c.ctas,
           c.threads, names[c.mode], c.sleep_ns, mean / rounds / (clock_khz * 1e-3), wall * 1e3, rounds);
    fflush(stdout);
  }
  return 0;
}
